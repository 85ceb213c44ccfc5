"""MJCF subset loader and closed-form constants for the quadrotor engine.

The reference never touches the physics itself: it loads ``model/drone/drone.xml``
with ``mujoco.MjModel.from_xml_path`` (reference: train_brax_ppo.py:202,
envs/hover_env.py:75) and calls ``mjx.step`` / ``mujoco.mj_step``.  This module
replaces that loader for the one model family the hot path needs:

    world -> base (free joint) -> 4 passive rotors (hinge about body z)
    + site-transmission motors on the base, + inertia-box fluid model.

``load_mjcf`` parses the subset of MJCF that affects dynamics into a generic
``TreeModel`` (what MuJoCo would compile: body tree, inertial frames, joints,
sites, motors, options).  ``derive_constants`` turns the tree into the constant
block the CUDA kernels integrate with (all derived in float64), and *verifies*
the structural assumptions of the closed form (balanced axisymmetric rotors,
spin axis = body z, ...), raising ``ModelError`` otherwise -- the reference
constructor may likewise raise on a bad XML.

Nothing here is on the per-step path; it runs once at engine creation.
"""
from __future__ import annotations

import math
import os
import xml.etree.ElementTree as ET
from dataclasses import dataclass, field

import numpy as np

__all__ = [
    "ModelError", "TreeModel", "QuadConstants", "load_mjcf", "derive_constants",
    "default_model_path", "load_default", "contact_report", "check_contacts",
]

JNT_FREE, JNT_HINGE = 0, 3   # MuJoCo's mjtJoint values for the two types we accept


class ModelError(ValueError):
    """The XML is outside the model family the engine integrates in closed form."""


# --------------------------------------------------------------------------
# small quaternion helpers (w, x, y, z), float64
# --------------------------------------------------------------------------
def quat_to_mat(q):
    w, x, y, z = (float(v) for v in q)
    n = math.sqrt(w * w + x * x + y * y + z * z)
    w, x, y, z = w / n, x / n, y / n, z / n
    return np.array([
        [1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
        [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
        [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)],
    ])


def _floats(s, n=None, default=None):
    if s is None:
        return None if default is None else np.array(default, dtype=np.float64)
    v = np.array([float(t) for t in s.replace(",", " ").split()], dtype=np.float64)
    if n is not None and v.size != n:
        raise ModelError(f"expected {n} numbers, got {v.size}: {s!r}")
    return v


# --------------------------------------------------------------------------
# generic tree (what the compiler would emit)
# --------------------------------------------------------------------------
@dataclass
class TreeModel:
    """Dynamics-relevant content of the MJCF, body 0 = world."""
    body_name: list = field(default_factory=list)
    body_parent: list = field(default_factory=list)
    body_pos: list = field(default_factory=list)      # frame in parent
    body_quat: list = field(default_factory=list)
    body_ipos: list = field(default_factory=list)     # inertial frame in body
    body_iquat: list = field(default_factory=list)
    body_mass: list = field(default_factory=list)
    body_inertia: list = field(default_factory=list)  # principal moments (3)
    # joints, in qpos order
    jnt_name: list = field(default_factory=list)
    jnt_type: list = field(default_factory=list)
    jnt_body: list = field(default_factory=list)
    jnt_pos: list = field(default_factory=list)
    jnt_axis: list = field(default_factory=list)
    jnt_damping: list = field(default_factory=list)
    jnt_armature: list = field(default_factory=list)
    # sites
    site_name: list = field(default_factory=list)
    site_body: list = field(default_factory=list)
    site_pos: list = field(default_factory=list)
    site_quat: list = field(default_factory=list)
    # motors (site transmission)
    act_name: list = field(default_factory=list)
    act_site: list = field(default_factory=list)
    act_gear: list = field(default_factory=list)      # 6
    act_ctrlrange: list = field(default_factory=list)  # (lo, hi) or None
    # collision geometry (never simulated: it only feeds the contact guard, ``contact_report``)
    geom_name: list = field(default_factory=list)
    geom_body: list = field(default_factory=list)
    geom_type: list = field(default_factory=list)      # plane | sphere | capsule | cylinder | ellipsoid | box | mesh
    geom_center: list = field(default_factory=list)    # centre of the bounding sphere, body frame (3)
    geom_rbound: list = field(default_factory=list)    # its radius (inf for a plane)
    geom_normal: list = field(default_factory=list)    # planes: unit normal in the body frame, else None
    geom_contype: list = field(default_factory=list)
    geom_conaffinity: list = field(default_factory=list)
    has_contact_section: bool = False                  # <contact> pairs / excludes are not interpreted
    # options
    timestep: float = 0.002
    gravity: np.ndarray = field(default_factory=lambda: np.array([0.0, 0.0, -9.81]))
    wind: np.ndarray = field(default_factory=lambda: np.zeros(3))
    density: float = 0.0
    viscosity: float = 0.0
    integrator: str = "Euler"

    @property
    def nbody(self):
        return len(self.body_name)

    @property
    def nq(self):
        return sum(7 if t == JNT_FREE else 1 for t in self.jnt_type)

    @property
    def nv(self):
        return sum(6 if t == JNT_FREE else 1 for t in self.jnt_type)

    @property
    def nu(self):
        return len(self.act_name)

    def qpos0(self):
        out = []
        for t, b in zip(self.jnt_type, self.jnt_body):
            if t == JNT_FREE:
                out += list(self.body_pos[b]) + list(self.body_quat[b])
            else:
                out.append(0.0)
        return np.array(out, dtype=np.float64)

    def summary(self):
        """Plain, order-stable view used to compare two parsed files."""
        def arr(x):
            return [None if v is None else np.asarray(v, dtype=np.float64).round(15).tolist() for v in x]
        return {
            "body_parent": list(self.body_parent), "body_pos": arr(self.body_pos),
            "body_quat": arr(self.body_quat), "body_ipos": arr(self.body_ipos),
            "body_iquat": arr(self.body_iquat), "body_mass": [float(m) for m in self.body_mass],
            "body_inertia": arr(self.body_inertia),
            "jnt_type": list(self.jnt_type), "jnt_body": list(self.jnt_body),
            "jnt_pos": arr(self.jnt_pos), "jnt_axis": arr(self.jnt_axis),
            "jnt_damping": list(self.jnt_damping), "jnt_armature": list(self.jnt_armature),
            "site_body": list(self.site_body), "site_pos": arr(self.site_pos), "site_quat": arr(self.site_quat),
            "act_site": list(self.act_site), "act_gear": arr(self.act_gear),
            "act_ctrlrange": arr(self.act_ctrlrange),
            "timestep": self.timestep, "gravity": self.gravity.tolist(), "wind": self.wind.tolist(),
            "density": self.density, "viscosity": self.viscosity, "integrator": self.integrator,
        }


def _collect_defaults(root):
    """class name -> {tag: attrib}, with nested <default> inheritance."""
    table = {}

    def walk(node, inherited, name):
        mine = {k: dict(v) for k, v in inherited.items()}
        for child in node:
            if child.tag != "default":
                mine.setdefault(child.tag, {}).update(child.attrib)
        table[name] = mine
        for child in node:
            if child.tag == "default":
                walk(child, mine, child.get("class", "main"))

    for d in root.findall("default"):
        walk(d, {}, d.get("class", "main"))
    table.setdefault("main", {})
    return table


def load_mjcf(path: str) -> TreeModel:
    """Parse the dynamics subset of an MJCF file (reference: model/drone/drone.xml).

    Handled: <compiler angle/autolimits>, <option timestep/gravity/wind/density/
    viscosity/integrator>, nested <default> classes for <motor>/<joint>, <body>
    (pos/quat), <inertial> (pos/quat/mass/diaginertia), <joint> (free|hinge; pos/
    axis/damping/armature), <site> (pos/quat), <actuator><motor site= gear=
    ctrlrange=>.  <geom> elements (incl. default classes, contype / conaffinity and the bounding radius of <mesh>
    assets read from their STL files) are parsed into bounding spheres for the contact guard (``contact_report``) --
    the engine has no collision stage, so a model whose geoms could touch must be rejected, not silently
    mis-simulated.  Sensors, visuals and <keyframe> are ignored (the reference's keyframe has the wrong length for
    nq and no code reads it).
    """
    try:
        root = ET.parse(path).getroot()
    except (ET.ParseError, OSError) as e:
        raise ModelError(f"cannot read MJCF {path!r}: {e}") from e
    if root.tag != "mujoco":
        raise ModelError("root element is not <mujoco>")
    m = TreeModel()
    defaults = _collect_defaults(root)

    comp = root.find("compiler")
    autolimits = True   # MuJoCo >= 2.3 default
    meshdir = ""
    if comp is not None:
        if comp.get("angle", "degree") not in ("degree", "radian"):
            raise ModelError("bad compiler angle")
        autolimits = comp.get("autolimits", "true") == "true"
        meshdir = comp.get("meshdir", comp.get("assetdir", ""))
    m.has_contact_section = root.find("contact") is not None
    meshes = {}
    for asset in root.findall("asset"):
        for me in asset.findall("mesh"):
            cls = me.get("class") or "main"
            d = defaults.get(cls, {}).get("mesh", {})
            fname = me.get("file")
            name = me.get("name") or (os.path.splitext(os.path.basename(fname))[0] if fname else None)
            meshes[name] = (None if fname is None else os.path.join(os.path.dirname(os.path.abspath(path)), meshdir, fname),
                            _floats(me.get("scale", d.get("scale")), 3, [1, 1, 1]))
    opt = root.find("option")
    if opt is not None:
        m.timestep = float(opt.get("timestep", m.timestep))
        m.gravity = _floats(opt.get("gravity"), 3, m.gravity)
        m.wind = _floats(opt.get("wind"), 3, m.wind)
        m.density = float(opt.get("density", 0.0))
        m.viscosity = float(opt.get("viscosity", 0.0))
        m.integrator = opt.get("integrator", "Euler")

    def dflt(tag, elem, childclass):
        cls = elem.get("class") or childclass or "main"
        if cls not in defaults:
            raise ModelError(f"unknown default class {cls!r}")
        return defaults[cls].get(tag, {})

    # world body
    m.body_name.append("world"); m.body_parent.append(0)
    m.body_pos.append(np.zeros(3)); m.body_quat.append(np.array([1.0, 0, 0, 0]))
    m.body_ipos.append(np.zeros(3)); m.body_iquat.append(np.array([1.0, 0, 0, 0]))
    m.body_mass.append(0.0); m.body_inertia.append(np.zeros(3))

    def add_geoms(elem, bid, childclass):
        for g in elem.findall("geom"):
            d = dflt("geom", g, childclass)
            get = lambda k, dv=None: g.get(k, d.get(k, dv))
            gtype = get("type", "mesh" if get("mesh") is not None else "sphere")
            pos = _floats(get("pos"), 3, [0, 0, 0])
            R = quat_to_mat(_floats(get("quat"), 4, [1, 0, 0, 0]))
            if get("fromto") is not None:
                ft = _floats(get("fromto"), 6)
                pos = 0.5 * (ft[:3] + ft[3:])
            size = _floats(get("size"), None, [0.0])
            normal = None
            if gtype == "plane":
                centre, rb, normal = pos, math.inf, R[:, 2]
            elif gtype == "sphere":
                centre, rb = pos, float(size[0])
            elif gtype in ("capsule", "cylinder"):
                half = 0.5 * float(np.linalg.norm(ft[3:] - ft[:3])) if get("fromto") is not None else float(size[1])
                centre, rb = pos, (float(size[0]) + half if gtype == "capsule" else math.hypot(float(size[0]), half))
            elif gtype == "ellipsoid":
                centre, rb = pos, float(np.max(size))
            elif gtype == "box":
                centre, rb = pos, float(np.linalg.norm(size[:3]))
            elif gtype == "mesh":
                mname = get("mesh")
                if mname not in meshes:
                    raise ModelError(f"geom references unknown mesh {mname!r}")
                v = _stl_vertices(*meshes[mname])
                if v is None:
                    centre, rb = pos, math.nan          # extent unknown: the contact guard refuses to clear it
                else:
                    w = v @ R.T + pos
                    centre = 0.5 * (w.min(axis=0) + w.max(axis=0))
                    rb = float(np.linalg.norm(w - centre, axis=1).max())
            else:
                raise ModelError(f"geom type {gtype!r} not supported")
            m.geom_name.append(g.get("name", f"geom{len(m.geom_name)}"))
            m.geom_body.append(bid); m.geom_type.append(gtype)
            m.geom_center.append(np.asarray(centre, dtype=np.float64)); m.geom_rbound.append(rb)
            m.geom_normal.append(normal)
            m.geom_contype.append(int(get("contype", 1))); m.geom_conaffinity.append(int(get("conaffinity", 1)))

    def add_sites(elem, bid, childclass):
        for s in elem.findall("site"):
            m.site_name.append(s.get("name", f"site{len(m.site_name)}"))
            m.site_body.append(bid)
            m.site_pos.append(_floats(s.get("pos"), 3, [0, 0, 0]))
            m.site_quat.append(_floats(s.get("quat"), 4, [1, 0, 0, 0]))

    def walk(elem, parent, childclass):
        for b in elem.findall("body"):
            bid = len(m.body_name)
            cc = b.get("childclass", childclass)
            m.body_name.append(b.get("name", f"body{bid}"))
            m.body_parent.append(parent)
            m.body_pos.append(_floats(b.get("pos"), 3, [0, 0, 0]))
            m.body_quat.append(_floats(b.get("quat"), 4, [1, 0, 0, 0]))
            ine = b.find("inertial")
            if ine is None:
                # the reference's bodies all carry <inertial>; mesh-inferred inertia is out of scope
                raise ModelError(f"body {m.body_name[-1]!r} has no <inertial>")
            if ine.get("fullinertia") is not None:
                raise ModelError("fullinertia not supported; give diaginertia + quat")
            m.body_ipos.append(_floats(ine.get("pos"), 3, [0, 0, 0]))
            m.body_iquat.append(_floats(ine.get("quat"), 4, [1, 0, 0, 0]))
            m.body_mass.append(float(ine.get("mass")))
            m.body_inertia.append(_floats(ine.get("diaginertia"), 3))
            joints = list(b.findall("joint")) + ([b.find("freejoint")] if b.find("freejoint") is not None else [])
            for j in joints:
                d = dflt("joint", j, cc)
                jtype = "free" if j.tag == "freejoint" else j.get("type", d.get("type", "hinge"))
                if jtype == "free":
                    m.jnt_type.append(JNT_FREE)
                elif jtype == "hinge":
                    m.jnt_type.append(JNT_HINGE)
                else:
                    raise ModelError(f"joint type {jtype!r} not supported")
                m.jnt_name.append(j.get("name", f"joint{len(m.jnt_name)}"))
                m.jnt_body.append(bid)
                m.jnt_pos.append(_floats(j.get("pos", d.get("pos")), 3, [0, 0, 0]))
                m.jnt_axis.append(_floats(j.get("axis", d.get("axis")), 3, [0, 0, 1]))
                m.jnt_damping.append(float(j.get("damping", d.get("damping", 0.0))))
                m.jnt_armature.append(float(j.get("armature", d.get("armature", 0.0))))
                for bad in ("stiffness", "frictionloss", "springref"):
                    if float(j.get(bad, d.get(bad, 0.0))) != 0.0:
                        raise ModelError(f"joint {bad} not supported")
                if j.get("range", d.get("range")) is not None and (
                        j.get("limited", d.get("limited", "auto")) != "false"):
                    raise ModelError("joint limits not supported")
            add_sites(b, bid, cc)
            add_geoms(b, bid, cc)
            walk(b, bid, cc)

    wb = root.find("worldbody")
    if wb is None:
        raise ModelError("no <worldbody>")
    add_sites(wb, 0, None)
    add_geoms(wb, 0, None)
    walk(wb, 0, None)

    act = root.find("actuator")
    if act is not None:
        for a in act:
            if a.tag != "motor":
                raise ModelError(f"actuator <{a.tag}> not supported (motors only)")
            d = dflt("motor", a, None)
            site = a.get("site", d.get("site"))
            if site is None or site not in m.site_name:
                raise ModelError("only site-transmission motors are supported")
            m.act_name.append(a.get("name", f"motor{len(m.act_name)}"))
            m.act_site.append(m.site_name.index(site))
            gear = _floats(a.get("gear", d.get("gear")), None, [1, 0, 0, 0, 0, 0])
            g6 = np.zeros(6); g6[:gear.size] = gear
            m.act_gear.append(g6)
            cr = a.get("ctrlrange", d.get("ctrlrange"))
            limited = a.get("ctrllimited", d.get("ctrllimited", "auto"))
            if cr is not None and (limited == "true" or (limited == "auto" and autolimits)):
                m.act_ctrlrange.append(_floats(cr, 2))
            else:
                m.act_ctrlrange.append(None)
    return m


def _stl_vertices(path, scale):
    """Vertices [n, 3] (float64, scaled) of a binary or ASCII STL file, or None when it cannot be read."""
    if path is None:
        return None
    try:
        with open(path, "rb") as f:
            raw = f.read()
    except OSError:
        return None
    v = None
    if len(raw) >= 84:
        ntri = int.from_bytes(raw[80:84], "little")
        if 84 + 50 * ntri == len(raw) and ntri > 0:
            tri = np.frombuffer(raw, dtype=np.uint8, offset=84).reshape(ntri, 50)
            v = tri[:, 12:48].copy().view("<f4").reshape(-1, 3).astype(np.float64)
    if v is None:
        pts = [ln.split()[1:4] for ln in raw.decode("ascii", "ignore").splitlines() if ln.strip().startswith("vertex")]
        if not pts:
            return None
        v = np.array(pts, dtype=np.float64)
    return v * np.asarray(scale, dtype=np.float64)


# --------------------------------------------------------------------------
# contact guard
# --------------------------------------------------------------------------
def contact_report(tree: TreeModel, pos_lo=None, pos_hi=None, overshoot: float = 0.0):
    """Geom pairs of ``tree`` that MuJoCo's collision phase could ever find in contact while the base origin stays in
    the box [pos_lo, pos_hi] (the env's termination bounds; None = unbounded), as a list of human-readable strings.

    The sm_100a step kernel has no collision / constraint stage, so it is only valid for models
    where this list is empty.  Filtering follows MuJoCo's broad phase: a pair is a candidate when
    ``(contype1 & conaffinity2) | (contype2 & conaffinity1)`` is non-zero, the geoms sit on different bodies, and the
    bodies are not parent and child (``filterparent``; the world body never counts as a parent).  The geometric test is
    conservative: every geom is replaced by a bounding sphere, a geom on a hinge body by the sphere swept about the
    hinge axis (valid for every joint angle), and the whole vehicle by one sphere about the base origin when tested
    against world geoms; ``overshoot`` inflates the box by the distance the base can travel in the one step that ends
    the episode.  A mesh whose file could not be read has an unknown extent and is reported.
    """
    out = []
    if tree.has_contact_section:
        out.append("<contact> section present: explicit pairs / excludes are not interpreted")
    ng = len(tree.geom_name)
    if ng == 0:
        return out
    free = [b for t, b in zip(tree.jnt_type, tree.jnt_body) if t == JNT_FREE]
    base = free[0] if free else None

    def to_base(b, x):
        """point x of body b's frame -> base frame (bodies are the base or its direct children, checked elsewhere)."""
        while b != base and b != 0:
            x = quat_to_mat(tree.body_quat[b]) @ x + tree.body_pos[b]
            b = tree.body_parent[b]
        return x, b

    # bounding spheres in the base frame (vehicle geoms) or the world frame (static geoms), valid for every hinge angle
    centre, radius, frame = [], [], []
    for g in range(ng):
        b = tree.geom_body[g]
        c = np.asarray(tree.geom_center[g], dtype=np.float64); r = tree.geom_rbound[g]
        hinge = [k for k, (t, jb) in enumerate(zip(tree.jnt_type, tree.jnt_body)) if t == JNT_HINGE and jb == b]
        if hinge and math.isfinite(r):
            k = hinge[0]
            ax = np.asarray(tree.jnt_axis[k], dtype=np.float64); ax = ax / np.linalg.norm(ax)
            rel = c - tree.jnt_pos[k]
            on_axis = tree.jnt_pos[k] + ax * float(rel @ ax)
            r = r + float(np.linalg.norm(c - on_axis))
            c = on_axis
        cb, root = to_base(b, c)
        centre.append(cb); radius.append(r); frame.append("world" if root == 0 and b == 0 else "vehicle")

    def candidates(i, j):
        bi, bj = tree.geom_body[i], tree.geom_body[j]
        if bi == bj:
            return False
        if not ((tree.geom_contype[i] & tree.geom_conaffinity[j]) or (tree.geom_contype[j] & tree.geom_conaffinity[i])):
            return False
        if bi != 0 and bj != 0 and (tree.body_parent[bi] == bj or tree.body_parent[bj] == bi):
            return False                                   # filterparent
        return True

    veh = [g for g in range(ng) if frame[g] == "vehicle"]
    R_vehicle = max([float(np.linalg.norm(centre[g])) + radius[g] for g in veh], default=0.0)
    lo = None if pos_lo is None else np.asarray(pos_lo, dtype=np.float64) - overshoot
    hi = None if pos_hi is None else np.asarray(pos_hi, dtype=np.float64) + overshoot
    for i in range(ng):
        for j in range(i + 1, ng):
            if not candidates(i, j):
                continue
            name = f"{tree.geom_name[i]} ({tree.body_name[tree.geom_body[i]]}) <-> {tree.geom_name[j]} ({tree.body_name[tree.geom_body[j]]})"
            if math.isnan(radius[i]) or math.isnan(radius[j]):
                out.append(name + ": mesh extent unknown (file not readable)")
                continue
            if frame[i] == frame[j] == "vehicle":
                gap = float(np.linalg.norm(centre[i] - centre[j])) - radius[i] - radius[j]
                if gap <= 0.0:
                    out.append(name + f": bounding spheres overlap by {-gap:.4f} m")
                continue
            if frame[i] == frame[j]:
                continue                                   # two static geoms
            w = i if frame[i] == "world" else j            # static geom vs the vehicle anywhere in the box
            if lo is None or hi is None or not (np.isfinite(lo).all() and np.isfinite(hi).all()):
                out.append(name + ": the env has no position bounds, a static geom is always reachable")
                continue
            if tree.geom_type[w] == "plane":
                n = np.asarray(tree.geom_normal[w], dtype=np.float64)
                dmin = float(np.minimum(n * lo, n * hi).sum() - n @ centre[w])    # lowest height of the base origin above the plane
                if dmin - R_vehicle <= 0.0:
                    out.append(name + f": plane reachable inside the termination box (clearance {dmin - R_vehicle:.4f} m)")
            else:
                d = np.maximum(np.maximum(lo - centre[w], centre[w] - hi), 0.0)
                gap = float(np.linalg.norm(d)) - radius[w] - R_vehicle
                if gap <= 0.0:
                    out.append(name + f": static geom reachable inside the termination box (clearance {gap:.4f} m)")
    return out


def check_contacts(tree: TreeModel, pos_lo=None, pos_hi=None, overshoot: float = 0.0, assume_no_contact: bool = False):
    """Raise ``ModelError`` when ``contact_report`` is non-empty, unless the caller explicitly takes responsibility
    with ``assume_no_contact=True``."""
    if assume_no_contact:
        return
    rep = contact_report(tree, pos_lo, pos_hi, overshoot)
    if rep:
        raise ModelError("the model has reachable contacts, which the engine does not simulate (pass "
                         "assume_no_contact=True to override):\n  " + "\n  ".join(rep))


# --------------------------------------------------------------------------
# closed-form constants for the fused kernels
# --------------------------------------------------------------------------
@dataclass
class QuadConstants:
    """Constants of the 10-DoF gyrostat + fluid closed form (float64).

    Coordinates: q = [p_O (world), quat wxyz, theta(4)], v = [dp_O (world), omega (body), s(4)].
    See DESIGN.md "Dynamics closed form" for the equations these feed.
    """
    dt: float
    gz: float                    # gravity z component (world); x, y must be 0
    mass: float                  # total
    com: np.ndarray              # composite COM in base frame (3)
    I_C: np.ndarray              # composite inertia about the COM, base axes (3x3)
    Ieff_inv: np.ndarray         # inv(I_C - zz^T sum J_i^2 / Js_i)
    rotor_J: np.ndarray          # axial inertia per rotor (4)
    rotor_Js: np.ndarray         # J + armature + dt*damping (4)
    rotor_damping: np.ndarray    # (4)
    rotor_r: np.ndarray          # rotor COM in base frame (4x3)
    wrench: np.ndarray           # (6x4): [force_body; torque about COM] per unit motor force
    ctrl_lo: np.ndarray          # (4) motor clamp
    ctrl_hi: np.ndarray
    # fluid, base body (inertial frame == body frame, at the origin)
    base_lin_visc: float
    base_lin_quad: np.ndarray    # (3) 0.5*rho*area_k
    base_ang_visc: float
    base_ang_quad: np.ndarray    # (3)
    base_box: np.ndarray
    # fluid, rotors (identical geometry is NOT assumed: per rotor)
    rot_lin_visc: np.ndarray     # (4)
    rot_lin_quad_ax: np.ndarray  # (4) axial  (along spin axis)
    rot_lin_quad_lat: np.ndarray  # (4) lateral (component-wise along the two spinning axes)
    rot_ang_visc: np.ndarray
    rot_ang_quad_ax: np.ndarray
    rot_ang_quad_lat: np.ndarray
    rot_box: np.ndarray          # (4x3)
    hover_thrust_per_motor: float


def _inertia_box(mass, inertia):
    """MuJoCo's equivalent inertia box (mj_inertiaBoxFluidModel), float64."""
    I0, I1, I2 = (float(v) for v in inertia)
    return np.array([
        math.sqrt(max(1e-15, I1 + I2 - I0) / mass * 6.0),
        math.sqrt(max(1e-15, I0 + I2 - I1) / mass * 6.0),
        math.sqrt(max(1e-15, I0 + I1 - I2) / mass * 6.0),
    ])


def _fluid_coeffs(box, rho, mu):
    diam = float(np.mean(box))
    lin_visc = 3.0 * math.pi * diam * mu
    ang_visc = math.pi * diam ** 3 * mu
    b0, b1, b2 = box
    lin_quad = 0.5 * rho * np.array([b1 * b2, b0 * b2, b0 * b1])
    ang_quad = rho / 64.0 * np.array([b0 * (b1 ** 4 + b2 ** 4), b1 * (b0 ** 4 + b2 ** 4), b2 * (b0 ** 4 + b1 ** 4)])
    return lin_visc, lin_quad, ang_visc, ang_quad


def derive_constants(m: TreeModel, tol: float = 1e-9) -> QuadConstants:
    """Check the tree is a base + 4 balanced axisymmetric z-rotors and derive the constants."""
    if m.integrator != "Euler":
        raise ModelError("only the Euler integrator is reproduced")
    if m.nbody != 6 or m.jnt_type != [JNT_FREE] + [JNT_HINGE] * 4:
        raise ModelError("expected world + free base + 4 hinge rotors")
    if m.body_parent[1] != 0 or any(p != 1 for p in m.body_parent[2:]):
        raise ModelError("rotors must be children of the free base")
    if m.jnt_body != [1, 2, 3, 4, 5]:
        raise ModelError("one joint per body expected, in body order")
    if m.jnt_damping[0] != 0.0 or m.jnt_armature[0] != 0.0:
        raise ModelError("damping/armature on the free joint not supported")
    if np.any(np.abs(m.gravity[:2]) > 0) or np.any(m.wind != 0):
        raise ModelError("gravity must be along z and wind must be zero")
    if np.linalg.norm(m.body_ipos[1]) > tol or np.linalg.norm(quat_to_mat(m.body_iquat[1]) - np.eye(3)) > tol:
        raise ModelError("base inertial frame must coincide with the base frame")
    if m.nu != 4:
        raise ModelError("expected 4 motors")

    dt = m.timestep
    mass = float(sum(m.body_mass))
    com = np.zeros(3)
    I_O = np.diag(np.asarray(m.body_inertia[1], dtype=np.float64)).copy()
    rotor_J = np.zeros(4); rotor_r = np.zeros((4, 3)); rot_box = np.zeros((4, 3))
    damping = np.zeros(4); armature = np.zeros(4)
    for k in range(4):
        b = 2 + k
        Rb = quat_to_mat(m.body_quat[b])            # rotor body frame in base frame (theta = 0)
        axis = Rb @ (m.jnt_axis[k + 1] / np.linalg.norm(m.jnt_axis[k + 1]))
        if np.linalg.norm(axis - np.array([0, 0, 1.0])) > tol:
            raise ModelError("rotor hinge axis must be the base z axis")
        anchor = m.body_pos[b] + Rb @ m.jnt_pos[k + 1]
        r = m.body_pos[b] + Rb @ m.body_ipos[b]
        if np.linalg.norm(np.cross(r - anchor, axis)) > tol:
            raise ModelError("rotor COM must lie on its spin axis (balanced rotor)")
        Ri = Rb @ quat_to_mat(m.body_iquat[b])       # inertial frame in base frame
        Ib = Ri @ np.diag(m.body_inertia[b]) @ Ri.T
        # axisymmetric about z: Ib = diag(Ia, Ia, J), no coupling
        J = Ib[2, 2]
        if (abs(Ib[0, 0] - Ib[1, 1]) > 1e-6 * J or
                np.max(np.abs(Ib - np.diag(np.diag(Ib)))) > 1e-6 * J):
            raise ModelError("rotor inertia must be axisymmetric about the spin axis")
        # which principal axis is the spin axis; the two lateral box sides must then be equal and the
        # lateral principal axes must sit at a multiple of 90 deg from base x at theta = 0
        ax = int(np.argmax(np.abs(Ri.T @ axis)))
        lat = [i for i in range(3) if i != ax]
        box = _inertia_box(m.body_mass[b], m.body_inertia[b])
        if abs(box[lat[0]] - box[lat[1]]) > 1e-9 * box[lat[0]]:
            raise ModelError("rotor lateral box sides differ")
        e = Ri[:, lat[0]]
        if min(abs(e[0]), abs(e[1])) > tol or abs(e[2]) > tol:
            raise ModelError("rotor lateral principal axes must be aligned with base x/y at theta=0")
        rot_box[k] = [box[ax], box[lat[0]], box[lat[1]]]   # (axial, lateral, lateral)
        rotor_J[k] = J
        rotor_r[k] = r
        damping[k] = m.jnt_damping[k + 1]
        armature[k] = m.jnt_armature[k + 1]
        mb = m.body_mass[b]
        com += mb * r
        I_O += Ib + mb * (np.dot(r, r) * np.eye(3) - np.outer(r, r))
    com /= mass
    I_C = I_O - mass * (np.dot(com, com) * np.eye(3) - np.outer(com, com))
    Js = rotor_J + armature + dt * damping
    zz = np.zeros((3, 3)); zz[2, 2] = 1.0
    Ieff = I_C - zz * float(np.sum(rotor_J ** 2 / Js))
    Ieff_inv = np.linalg.inv(Ieff)

    # motors: site wrench in base frame, torque referred to the composite COM
    wrench = np.zeros((6, 4)); lo = np.zeros(4); hi = np.zeros(4)
    for k in range(4):
        s = m.act_site[k]
        if m.site_body[s] != 1:
            raise ModelError("thrust sites must belong to the base body")
        Rs = quat_to_mat(m.site_quat[s])
        f = Rs @ m.act_gear[k][:3]
        t = Rs @ m.act_gear[k][3:]
        wrench[:3, k] = f
        wrench[3:, k] = t + np.cross(m.site_pos[s] - com, f)
        if m.act_ctrlrange[k] is None:
            lo[k], hi[k] = -np.inf, np.inf
        else:
            lo[k], hi[k] = m.act_ctrlrange[k]

    rho, mu = m.density, m.viscosity
    base_box = _inertia_box(m.body_mass[1], m.body_inertia[1])
    blv, blq, bav, baq = _fluid_coeffs(base_box, rho, mu)
    rlv = np.zeros(4); rlqa = np.zeros(4); rlql = np.zeros(4)
    rav = np.zeros(4); raqa = np.zeros(4); raql = np.zeros(4)
    for k in range(4):
        lv, lq, av, aq = _fluid_coeffs(rot_box[k], rho, mu)   # box ordered (axial, lat, lat)
        rlv[k], rlqa[k], rlql[k] = lv, lq[0], lq[1]
        rav[k], raqa[k], raql[k] = av, aq[0], aq[1]

    fz = float(np.sum(wrench[2]))
    hover = mass * abs(m.gravity[2]) / fz if fz > 0 else float("nan")
    return QuadConstants(
        dt=dt, gz=float(m.gravity[2]), mass=mass, com=com, I_C=I_C, Ieff_inv=Ieff_inv,
        rotor_J=rotor_J, rotor_Js=Js, rotor_damping=damping, rotor_r=rotor_r, wrench=wrench,
        ctrl_lo=lo, ctrl_hi=hi,
        base_lin_visc=blv, base_lin_quad=blq, base_ang_visc=bav, base_ang_quad=baq, base_box=base_box,
        rot_lin_visc=rlv, rot_lin_quad_ax=rlqa, rot_lin_quad_lat=rlql,
        rot_ang_visc=rav, rot_ang_quad_ax=raqa, rot_ang_quad_lat=raql, rot_box=rot_box,
        hover_thrust_per_motor=hover,
    )


def default_model_path() -> str:
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "assets", "quad_x2.xml")


def load_default():
    tree = load_mjcf(default_model_path())
    return tree, derive_constants(tree)
