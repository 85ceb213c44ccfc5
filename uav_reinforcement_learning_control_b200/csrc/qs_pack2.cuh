// qs_pack2.cuh -- two float32 lanes per register pair on Blackwell's packed FP32 pipe (device code only).
//
// sm_100a executes add / mul / fma on .f32x2 operands (SASS FADD2 / FMUL2 / FFMA2): one issue slot, two IEEE-754
// round-to-nearest results.  ptxas folds negation, absolute value, scalar broadcast (a uniform register or an
// immediate) into the packed instruction's operand modifiers, so a scalar constant costs nothing extra.  `f2` carries
// the SAME quantity of TWO envs; every operator below is lane-wise and exactly the scalar operation of qs_math.cuh on
// each half, which lets the templated dynamics (qs_dynamics.cuh: physics_step_t<R>) serve both R = float and R = f2
// from one source.  Used by the packed step kernel (qs_step2.cuh), which is issue-bound, not FP32-pipe-bound.
#pragma once

#if defined(__CUDACC__)

namespace qs {

struct f2 { float x, y; };                      // x: even env of the pair, y: odd env
struct b2 { bool x, y; };                       // lane-wise predicate

typedef unsigned long long u64_;
__device__ __forceinline__ u64_ pk_(f2 a) { u64_ r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a.x), "f"(a.y)); return r; }
__device__ __forceinline__ f2 upk_(u64_ v) { f2 r; asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v)); return r; }
__device__ __forceinline__ f2 bc_(float c) { return f2{c, c}; }
__device__ __forceinline__ f2 bc_(f2 c) { return c; }
template <> __host__ __device__ __forceinline__ f2 splat_<f2>(float c) { return f2{c, c}; }

__device__ __forceinline__ f2 operator+(f2 a, f2 b) { u64_ d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(pk_(a)), "l"(pk_(b))); return upk_(d); }
__device__ __forceinline__ f2 operator-(f2 a, f2 b) { u64_ d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(pk_(a)), "l"(pk_(b))); return upk_(d); }
__device__ __forceinline__ f2 operator*(f2 a, f2 b) { u64_ d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(pk_(a)), "l"(pk_(b))); return upk_(d); }
__device__ __forceinline__ f2 operator-(f2 a) { return f2{-a.x, -a.y}; }                 // folded into the consumer's operand modifier
__device__ __forceinline__ f2 operator+(f2 a, float b) { return a + bc_(b); }
__device__ __forceinline__ f2 operator+(float a, f2 b) { return bc_(a) + b; }
__device__ __forceinline__ f2 operator-(f2 a, float b) { return a - bc_(b); }
__device__ __forceinline__ f2 operator-(float a, f2 b) { return bc_(a) - b; }
__device__ __forceinline__ f2 operator*(f2 a, float b) { return a * bc_(b); }
__device__ __forceinline__ f2 operator*(float a, f2 b) { return bc_(a) * b; }
__device__ __forceinline__ f2& operator+=(f2& a, f2 b) { a = a + b; return a; }
__device__ __forceinline__ f2& operator-=(f2& a, f2 b) { a = a - b; return a; }

__device__ __forceinline__ f2 fma2_(f2 a, f2 b, f2 c) {
    u64_ d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(pk_(a)), "l"(pk_(b)), "l"(pk_(c)));
    return upk_(d);
}
__device__ __forceinline__ f2 fma_(f2 a, f2 b, f2 c) { return fma2_(a, b, c); }
__device__ __forceinline__ f2 fma_(float a, f2 b, f2 c) { return fma2_(bc_(a), b, c); }
__device__ __forceinline__ f2 fma_(f2 a, float b, f2 c) { return fma2_(a, bc_(b), c); }
__device__ __forceinline__ f2 fma_(f2 a, f2 b, float c) { return fma2_(a, b, bc_(c)); }
__device__ __forceinline__ f2 fma_(float a, f2 b, float c) { return fma2_(bc_(a), b, bc_(c)); }
__device__ __forceinline__ f2 fma_(f2 a, float b, float c) { return fma2_(a, bc_(b), bc_(c)); }
__device__ __forceinline__ f2 fma_(float a, float b, f2 c) { return fma2_(bc_(a), bc_(b), c); }

// lane-wise forms of the scalar helpers of qs_math.cuh (MUFU / min-max / compare have no packed instruction)
__device__ __forceinline__ f2 abs_(f2 a) { return f2{fabsf(a.x), fabsf(a.y)}; }         // folded into |R| operand modifiers
__device__ __forceinline__ f2 min_(f2 a, f2 b) { return f2{fminf(a.x, b.x), fminf(a.y, b.y)}; }
__device__ __forceinline__ f2 max_(f2 a, f2 b) { return f2{fmaxf(a.x, b.x), fmaxf(a.y, b.y)}; }
__device__ __forceinline__ f2 clamp_(f2 a, float lo, float hi) { return f2{clamp_(a.x, lo, hi), clamp_(a.y, lo, hi)}; }
__device__ __forceinline__ f2 rsqrt_(f2 a) { return f2{rsqrt_(a.x), rsqrt_(a.y)}; }
__device__ __forceinline__ f2 rcp_(f2 a) { return f2{rcp_(a.x), rcp_(a.y)}; }
__device__ __forceinline__ f2 sqrt_(f2 a) { return f2{sqrt_(a.x), sqrt_(a.y)}; }
__device__ __forceinline__ f2 exp_(f2 a) { return f2{exp_(a.x), exp_(a.y)}; }
__device__ __forceinline__ f2 copysign_(f2 a, f2 s) { return f2{copysignf(a.x, s.x), copysignf(a.y, s.y)}; }
__device__ __forceinline__ f2 sel_(b2 m, f2 a, f2 b) { return f2{m.x ? a.x : b.x, m.y ? a.y : b.y}; }
__device__ __forceinline__ f2 sel_(b2 m, f2 a, float b) { return f2{m.x ? a.x : b, m.y ? a.y : b}; }
__device__ __forceinline__ b2 gt_(f2 a, float b) { return b2{a.x > b, a.y > b}; }
__device__ __forceinline__ b2 gt_(f2 a, f2 b) { return b2{a.x > b.x, a.y > b.y}; }
__device__ __forceinline__ b2 lt_(f2 a, float b) { return b2{a.x < b, a.y < b}; }
__device__ __forceinline__ b2 ge_(f2 a, float b) { return b2{a.x >= b, a.y >= b}; }
__device__ __forceinline__ b2 le_(f2 a, float b) { return b2{a.x <= b, a.y <= b}; }
__device__ __forceinline__ b2 isnan_(f2 a) { return b2{a.x != a.x, a.y != a.y}; }
__device__ __forceinline__ b2 operator&&(b2 a, b2 b) { return b2{a.x && b.x, a.y && b.y}; }
__device__ __forceinline__ b2 operator||(b2 a, b2 b) { return b2{a.x || b.x, a.y || b.y}; }

__device__ __forceinline__ void sincos_fast_(f2 a, f2* s, f2* c) {
    sincos_fast_(a.x, &s->x, &c->x);
    sincos_fast_(a.y, &s->y, &c->y);
}

// accurate sin / cos: the Taylor pair of qs_math.cuh: sincos_ on both lanes at once when both arguments are inside its
// range (always, on the hot path: |w| dt / 2 <= 0.5 up to |w| = 100 rad/s); lane-wise scalar calls otherwise
__device__ __forceinline__ void sincos_(f2 x, f2* s, f2* c) {
    if (fabsf(x.x) <= 0.5f && fabsf(x.y) <= 0.5f) {
        const f2 z = x * x;
        f2 ps = bc_(-2.5052108385441720e-08f);
        ps = fma_(ps, z, 2.7557319223985893e-06f);
        ps = fma_(ps, z, -1.9841269841269841e-04f);
        ps = fma_(ps, z, 8.3333333333333332e-03f);
        ps = fma_(ps, z, -1.6666666666666666e-01f);
        f2 pc = bc_(-2.7557319223985888e-07f);
        pc = fma_(pc, z, 2.4801587301587302e-05f);
        pc = fma_(pc, z, -1.3888888888888889e-03f);
        pc = fma_(pc, z, 4.1666666666666664e-02f);
        pc = fma_(pc, z, -0.5f);
        *s = fma_(x * z, ps, x);
        *c = fma_(z, pc, 1.0f);
        return;
    }
    sincos_(x.x, &s->x, &c->x);
    sincos_(x.y, &s->y, &c->y);
}

// atan2 / asin of qs_math.cuh, lane-wise around the shared (packed) polynomial
__device__ __forceinline__ f2 atan2_(f2 y, f2 x) {
    const f2 ax = abs_(x), ay = abs_(y);
    const f2 mx = max_(ax, ay), mn = min_(ax, ay);
    const f2 t = sel_(gt_(mx, 1e-30f), mn * rcp_(mx), 0.f);
    f2 r = atan_unit_(t);
    r = sel_(gt_(ay, ax), 1.5707963267948966f - r, r);
    r = sel_(lt_(x, 0.f), 3.141592653589793f - r, r);
    r = copysign_(r, y);
    const f2 chk = x + y;
    return sel_(isnan_(chk), chk, r);
}

__device__ __forceinline__ f2 asin_unit_(f2 x) {
    const f2 a = fma_(-x, x, 1.0f);
    f2 c;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(c.x) : "f"(a.x));
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(c.y) : "f"(a.y));
    return 2.0f * atan_unit_(x * rcp_(1.0f + c));
}

}  // namespace qs

#endif  // __CUDACC__
