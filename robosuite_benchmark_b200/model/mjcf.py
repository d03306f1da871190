"""MJCF-subset compiler: XML -> flat model arrays (the fields of include/rsb_model.h).

This stands where the reference relies on mujoco-py's `load_model_from_xml` + `MjSim`
(reached through `suite.make`, reference util/rlkit_utils.py:49-56).  MuJoCo is not available
in the build or run environment, so the compile-time quantities MuJoCo's
compiler / mj_setConst would produce are computed here: tree order,
dof parent chains and sparse-M addresses, inertial frames inferred from geoms, bounding radii,
mean inertia, `dof_invweight0` / `body_invweight0`, and the statically filtered candidate
contact pairs in MuJoCo's contact order.

Supported subset: <compiler angle>, <option>, <worldbody>/<body>/<inertial>/<joint>/<freejoint>/
<geom>/<site>, <actuator><motor>/<position>, <contact><exclude>.  Joints: free, hinge, slide.
Geoms: plane, sphere, capsule, cylinder, box.
"""
from __future__ import annotations

import xml.etree.ElementTree as ET
from dataclasses import dataclass, field
from typing import Dict, List

import numpy as np

JNT_FREE, JNT_SLIDE, JNT_HINGE = 0, 2, 3
GEOM_PLANE, GEOM_SPHERE, GEOM_CAPSULE, GEOM_CYLINDER, GEOM_BOX = 0, 2, 3, 5, 6
_GEOM_TYPES = {"plane": GEOM_PLANE, "sphere": GEOM_SPHERE, "capsule": GEOM_CAPSULE,
               "cylinder": GEOM_CYLINDER, "box": GEOM_BOX}
_JNT_TYPES = {"free": JNT_FREE, "slide": JNT_SLIDE, "hinge": JNT_HINGE}

MINVAL = 1e-15


# ----------------------------------------------------------------------------- quaternion helpers
def quat_mul(a, b):
    w1, x1, y1, z1 = a
    w2, x2, y2, z2 = b
    return np.array([w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2,
                     w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2,
                     w1 * y2 - x1 * z2 + y1 * w2 + z1 * x2,
                     w1 * z2 + x1 * y2 - y1 * x2 + z1 * w2])


def quat2mat(q):
    w, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def mat2quat(R):
    """Rotation matrix -> unit quaternion (w,x,y,z), w >= 0 branch preferred."""
    t = np.trace(R)
    if t > 0:
        s = np.sqrt(t + 1.0) * 2
        q = np.array([0.25 * s, (R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s])
    else:
        i = int(np.argmax(np.diag(R)))
        j, k = (i + 1) % 3, (i + 2) % 3
        s = np.sqrt(1.0 + R[i, i] - R[j, j] - R[k, k]) * 2
        q = np.zeros(4)
        q[0] = (R[k, j] - R[j, k]) / s
        q[1 + i] = 0.25 * s
        q[1 + j] = (R[j, i] + R[i, j]) / s
        q[1 + k] = (R[k, i] + R[i, k]) / s
    q /= np.linalg.norm(q)
    if q[0] < 0:
        q = -q
    return q


def axisangle2quat(axis, angle):
    axis = np.asarray(axis, float)
    axis = axis / np.linalg.norm(axis)
    return np.concatenate([[np.cos(angle / 2)], np.sin(angle / 2) * axis])


def _floats(s, n=None, default=None):
    if s is None:
        return None if default is None else np.array(default, float)
    v = np.array([float(x) for x in s.split()], float)
    if n is not None and v.size != n:
        if v.size < n and default is not None:
            out = np.array(default, float)
            out[:v.size] = v
            return out
        raise ValueError(f"expected {n} numbers, got {s!r}")
    return v


# ----------------------------------------------------------------------------- compiled model
@dataclass
class Model:
    """Flat arrays, MuJoCo naming.  See include/rsb_model.h for the C view."""
    opt: Dict[str, float] = field(default_factory=dict)
    names: Dict[str, Dict[str, int]] = field(default_factory=dict)  # kind -> name -> id
    arrays: Dict[str, np.ndarray] = field(default_factory=dict)

    def __getattr__(self, k):
        arr = self.__dict__.get("arrays", {})
        if k in arr:
            return arr[k]
        opt = self.__dict__.get("opt", {})
        if k in opt:
            return opt[k]
        raise AttributeError(k)

    def id(self, kind: str, name: str) -> int:
        return self.names[kind][name]


# ----------------------------------------------------------------------------- geometry / inertia
def _geom_volume_inertia(gtype, size):
    """Volume and unit-density inertia diag (about the geom centre, geom frame)."""
    if gtype == GEOM_BOX:
        a, b, c = size
        vol = 8 * a * b * c
        I = vol / 3.0 * np.array([b * b + c * c, a * a + c * c, a * a + b * b])
    elif gtype == GEOM_SPHERE:
        r = size[0]
        vol = 4.0 / 3.0 * np.pi * r ** 3
        I = np.full(3, 0.4 * vol * r * r)
    elif gtype == GEOM_CYLINDER:
        r, h = size[0], size[1]
        vol = np.pi * r * r * 2 * h
        ixy = vol * (3 * r * r + 4 * h * h) / 12.0
        I = np.array([ixy, ixy, 0.5 * vol * r * r])
    elif gtype == GEOM_CAPSULE:
        r, h = size[0], size[1]
        vc = np.pi * r * r * 2 * h
        vs = 4.0 / 3.0 * np.pi * r ** 3
        vol = vc + vs
        izz = 0.5 * vc * r * r + 0.4 * vs * r * r
        ixy = vc * (3 * r * r + 4 * h * h) / 12.0 + vs * (0.4 * r * r + h * h + 0.75 * r * h)
        I = np.array([ixy, ixy, izz])
    else:
        vol, I = 0.0, np.zeros(3)
    return vol, I


def _rbound(gtype, size):
    if gtype == GEOM_BOX:
        return float(np.linalg.norm(size))
    if gtype == GEOM_SPHERE:
        return float(size[0])
    if gtype in (GEOM_CAPSULE,):
        return float(size[0] + size[1])
    if gtype == GEOM_CYLINDER:
        return float(np.hypot(size[0], size[1]))
    return 0.0  # plane: unbounded, handled separately


# ----------------------------------------------------------------------------- compiler
def compile_mjcf(xml: str) -> Model:
    root = ET.fromstring(xml)
    comp = root.find("compiler")
    angle_deg = True
    if comp is not None and comp.get("angle", "degree") == "radian":
        angle_deg = False
    ang = (np.pi / 180.0) if angle_deg else 1.0

    # ---- options (MuJoCo defaults; robosuite base.xml overrides come from the XML itself)
    opt = dict(timestep=0.002, gravity=np.array([0, 0, -9.81]), impratio=1.0, tolerance=1e-8,
               ls_tolerance=0.01, cone=0, iterations=100, ls_iterations=50)
    o = root.find("option")
    if o is not None:
        for k in ("timestep", "impratio", "tolerance", "ls_tolerance"):
            if o.get(k) is not None:
                opt[k] = float(o.get(k))
        for k in ("iterations", "ls_iterations"):
            if o.get(k) is not None:
                opt[k] = int(o.get(k))
        if o.get("gravity") is not None:
            opt["gravity"] = _floats(o.get("gravity"), 3)
        if o.get("cone") is not None:
            opt["cone"] = {"pyramidal": 0, "elliptic": 1}[o.get("cone")]
        if o.get("integrator", "Euler") != "Euler":
            raise ValueError("only the Euler integrator is supported")
        if o.get("solver", "Newton") != "Newton":
            raise ValueError("only the Newton solver is supported")

    # ---- default geom/joint attributes (class-less <default> only)
    dgeom, djoint = {}, {}
    d = root.find("default")
    if d is not None:
        if d.find("geom") is not None:
            dgeom = dict(d.find("geom").attrib)
        if d.find("joint") is not None:
            djoint = dict(d.find("joint").attrib)

    bodies, joints, geoms, sites = [], [], [], []
    names = {"body": {}, "joint": {}, "geom": {}, "site": {}, "actuator": {}}

    def parse_pose(e):
        pos = _floats(e.get("pos"), 3, [0, 0, 0])
        if e.get("quat") is not None:
            quat = _floats(e.get("quat"), 4)
            quat = quat / np.linalg.norm(quat)
        elif e.get("axisangle") is not None:
            aa = _floats(e.get("axisangle"), 4)
            quat = axisangle2quat(aa[:3], aa[3] * ang)
        elif e.get("euler") is not None:
            ex, ey, ez = _floats(e.get("euler"), 3) * ang  # default eulerseq "xyz" (intrinsic)
            quat = quat_mul(quat_mul(axisangle2quat([1, 0, 0], ex), axisangle2quat([0, 1, 0], ey)),
                            axisangle2quat([0, 0, 1], ez))
        else:
            quat = np.array([1.0, 0, 0, 0])
        return pos, quat

    def add_body(e, parent):
        bid = len(bodies)
        name = e.get("name", f"body{bid}") if bid else "world"
        pos, quat = parse_pose(e) if bid else (np.zeros(3), np.array([1.0, 0, 0, 0]))
        b = dict(name=name, parent=parent, pos=pos, quat=quat, joints=[], geoms=[], inertial=None)
        bodies.append(b)
        names["body"][name] = bid
        for c in e:
            if c.tag == "inertial":
                ipos, iquat = parse_pose(c)
                if c.get("fullinertia") is not None:
                    raise ValueError("fullinertia unsupported")
                b["inertial"] = dict(pos=ipos, quat=iquat, mass=float(c.get("mass")),
                                     inertia=_floats(c.get("diaginertia"), 3))
            elif c.tag in ("joint", "freejoint"):
                a = dict(djoint)
                a.update(c.attrib)
                jt = JNT_FREE if c.tag == "freejoint" else _JNT_TYPES[a.get("type", "hinge")]
                rng = _floats(a.get("range"), 2, [0, 0])
                limited = a.get("limited")
                if limited is None or limited == "auto":
                    limited = a.get("range") is not None
                else:
                    limited = limited == "true"
                if jt == JNT_HINGE:
                    rng = rng * ang
                axis = _floats(a.get("axis"), 3, [0, 0, 1])
                axis = axis / max(np.linalg.norm(axis), MINVAL)
                j = dict(name=a.get("name", f"joint{len(joints)}"), type=jt, body=bid,
                         pos=_floats(a.get("pos"), 3, [0, 0, 0]), axis=axis, range=rng,
                         limited=bool(limited) and jt != JNT_FREE,
                         damping=float(a.get("damping", 0)), armature=float(a.get("armature", 0)),
                         frictionloss=float(a.get("frictionloss", 0)),
                         stiffness=float(a.get("stiffness", 0)),
                         springref=float(a.get("springref", 0)) * (ang if jt == JNT_HINGE else 1.0),
                         ref=float(a.get("ref", 0)) * (ang if jt == JNT_HINGE else 1.0),
                         margin=float(a.get("margin", 0)),
                         solref=_floats(a.get("solreflimit"), 2, [0.02, 1.0]),
                         solimp=_floats(a.get("solimplimit"), 5, [0.9, 0.95, 0.001, 0.5, 2.0]),
                         solref_fr=_floats(a.get("solreffriction"), 2, [0.02, 1.0]),
                         solimp_fr=_floats(a.get("solimpfriction"), 5, [0.9, 0.95, 0.001, 0.5, 2.0]))
                names["joint"][j["name"]] = len(joints)
                b["joints"].append(len(joints))
                joints.append(j)
            elif c.tag == "geom":
                a = dict(dgeom)
                a.update(c.attrib)
                gt = _GEOM_TYPES[a.get("type", "sphere")]
                size = _floats(a.get("size"), None, [0, 0, 0])
                size = np.concatenate([size, np.zeros(3 - size.size)]) if size.size < 3 else size[:3]
                gpos, gquat = parse_pose(c)
                if a.get("fromto") is not None:
                    ft = _floats(a.get("fromto"), 6)
                    p0, p1 = ft[:3], ft[3:]
                    gpos = 0.5 * (p0 + p1)
                    dvec = p1 - p0
                    ln = np.linalg.norm(dvec)
                    size = np.array([size[0], ln / 2, 0.0])
                    zax = dvec / ln
                    v = np.cross([0, 0, 1.0], zax)
                    s = np.linalg.norm(v)
                    if s < 1e-12:
                        gquat = np.array([1.0, 0, 0, 0]) if zax[2] > 0 else np.array([0, 1.0, 0, 0])
                    else:
                        gquat = axisangle2quat(v / s, np.arctan2(s, zax[2]))
                fr = _floats(a.get("friction"), 3, [1, 0.005, 0.0001])
                g = dict(name=a.get("name", f"geom{len(geoms)}"), type=gt, body=bid, size=size,
                         pos=gpos, quat=gquat, contype=int(a.get("contype", 1)),
                         conaffinity=int(a.get("conaffinity", 1)), condim=int(a.get("condim", 3)),
                         priority=int(a.get("priority", 0)), friction=fr,
                         solmix=float(a.get("solmix", 1)), solref=_floats(a.get("solref"), 2, [0.02, 1.0]),
                         solimp=_floats(a.get("solimp"), 5, [0.9, 0.95, 0.001, 0.5, 2.0]),
                         margin=float(a.get("margin", 0)), gap=float(a.get("gap", 0)),
                         density=float(a.get("density", 1000)),
                         mass=None if a.get("mass") is None else float(a.get("mass")),
                         group=int(a.get("group", 0)))
                names["geom"][g["name"]] = len(geoms)
                b["geoms"].append(len(geoms))
                geoms.append(g)
            elif c.tag == "site":
                spos, squat = parse_pose(c)
                s = dict(name=c.get("name", f"site{len(sites)}"), body=bid, pos=spos, quat=squat)
                names["site"][s["name"]] = len(sites)
                sites.append(s)
            elif c.tag == "body":
                add_body(c, bid)
        return bid

    add_body(root.find("worldbody"), 0)
    bodies[0]["parent"] = 0
    nbody, njnt, ngeom, nsite = len(bodies), len(joints), len(geoms), len(sites)

    # ---- address spaces
    qadr, dadr = 0, 0
    for j in joints:
        j["qposadr"], j["dofadr"] = qadr, dadr
        if j["type"] == JNT_FREE:
            qadr, dadr = qadr + 7, dadr + 6
        else:
            qadr, dadr = qadr + 1, dadr + 1
    nq, nv = qadr, dadr

    A: Dict[str, np.ndarray] = {}
    I32 = np.int32
    A["body_parentid"] = np.array([b["parent"] for b in bodies], I32)
    rootid = np.zeros(nbody, I32)
    for i in range(1, nbody):
        p = bodies[i]["parent"]
        rootid[i] = i if p == 0 else rootid[p]
    A["body_rootid"] = rootid
    A["body_jntnum"] = np.array([len(b["joints"]) for b in bodies], I32)
    A["body_jntadr"] = np.array([b["joints"][0] if b["joints"] else -1 for b in bodies], I32)
    A["body_pos"] = np.array([b["pos"] for b in bodies], float)
    A["body_quat"] = np.array([b["quat"] for b in bodies], float)

    # dofs
    dof_bodyid, dof_jntid = np.zeros(nv, I32), np.zeros(nv, I32)
    dof_arm, dof_damp, dof_fl = np.zeros(nv), np.zeros(nv), np.zeros(nv)
    dof_solref, dof_solimp = np.zeros((nv, 2)), np.zeros((nv, 5))
    for ji, j in enumerate(joints):
        n = 6 if j["type"] == JNT_FREE else 1
        sl = slice(j["dofadr"], j["dofadr"] + n)
        dof_bodyid[sl], dof_jntid[sl] = j["body"], ji
        dof_arm[sl], dof_damp[sl], dof_fl[sl] = j["armature"], j["damping"], j["frictionloss"]
        dof_solref[sl], dof_solimp[sl] = j["solref_fr"], j["solimp_fr"]
    body_dofnum, body_dofadr = np.zeros(nbody, I32), np.full(nbody, -1, I32)
    for bi, b in enumerate(bodies):
        n = sum(6 if joints[j]["type"] == JNT_FREE else 1 for j in b["joints"])
        body_dofnum[bi] = n
        if n:
            body_dofadr[bi] = joints[b["joints"][0]]["dofadr"]
    dof_parentid = np.full(nv, -1, I32)
    last_dof_of_body = np.full(nbody, -1, I32)  # last dof on the chain ending at this body
    for bi in range(1, nbody):
        p = bodies[bi]["parent"]
        last = last_dof_of_body[p]
        for k in range(body_dofnum[bi]):
            dof = body_dofadr[bi] + k
            dof_parentid[dof] = last
            last = dof
        last_dof_of_body[bi] = last
    dof_Madr = np.zeros(nv, I32)
    nM = 0
    for i in range(nv):
        dof_Madr[i] = nM
        k = i
        while k >= 0:
            nM += 1
            k = dof_parentid[k]
    A.update(body_dofnum=body_dofnum, body_dofadr=body_dofadr, dof_bodyid=dof_bodyid, dof_jntid=dof_jntid,
             dof_parentid=dof_parentid, dof_Madr=dof_Madr, dof_armature=dof_arm, dof_damping=dof_damp,
             dof_frictionloss=dof_fl, dof_solref=dof_solref, dof_solimp=dof_solimp)

    A["jnt_type"] = np.array([j["type"] for j in joints], I32)
    A["jnt_qposadr"] = np.array([j["qposadr"] for j in joints], I32)
    A["jnt_dofadr"] = np.array([j["dofadr"] for j in joints], I32)
    A["jnt_bodyid"] = np.array([j["body"] for j in joints], I32)
    A["jnt_limited"] = np.array([int(j["limited"]) for j in joints], I32)
    A["jnt_pos"] = np.array([j["pos"] for j in joints], float).reshape(njnt, 3)
    A["jnt_axis"] = np.array([j["axis"] for j in joints], float).reshape(njnt, 3)
    A["jnt_range"] = np.array([j["range"] for j in joints], float).reshape(njnt, 2)
    A["jnt_stiffness"] = np.array([j["stiffness"] for j in joints], float)
    A["jnt_margin"] = np.array([j["margin"] for j in joints], float)
    A["jnt_solref"] = np.array([j["solref"] for j in joints], float).reshape(njnt, 2)
    A["jnt_solimp"] = np.array([j["solimp"] for j in joints], float).reshape(njnt, 5)

    # qpos0 / qpos_spring
    qpos0, qspring = np.zeros(nq), np.zeros(nq)
    for j in joints:
        a = j["qposadr"]
        if j["type"] == JNT_FREE:
            b = bodies[j["body"]]
            qpos0[a:a + 3], qpos0[a + 3:a + 7] = b["pos"], b["quat"]
            qspring[a:a + 7] = qpos0[a:a + 7]
        else:
            qpos0[a], qspring[a] = j["ref"], j["springref"]
    A["qpos0"], A["qpos_spring"] = qpos0, qspring

    # geoms / sites
    A["geom_type"] = np.array([g["type"] for g in geoms], I32)
    A["geom_bodyid"] = np.array([g["body"] for g in geoms], I32)
    A["geom_size"] = np.array([g["size"] for g in geoms], float).reshape(ngeom, 3)
    A["geom_pos"] = np.array([g["pos"] for g in geoms], float).reshape(ngeom, 3)
    A["geom_quat"] = np.array([g["quat"] for g in geoms], float).reshape(ngeom, 4)
    A["geom_rbound"] = np.array([_rbound(g["type"], g["size"]) for g in geoms], float)
    A["geom_contype"] = np.array([g["contype"] for g in geoms], I32)
    A["geom_conaffinity"] = np.array([g["conaffinity"] for g in geoms], I32)
    A["site_bodyid"] = np.array([s["body"] for s in sites], I32)
    A["site_pos"] = np.array([s["pos"] for s in sites], float).reshape(nsite, 3)
    A["site_quat"] = np.array([s["quat"] for s in sites], float).reshape(nsite, 4)

    # ---- body inertial frames (explicit <inertial> wins; else inferred from geoms)
    body_mass, body_inertia = np.zeros(nbody), np.zeros((nbody, 3))
    body_ipos, body_iquat = np.zeros((nbody, 3)), np.tile([1.0, 0, 0, 0], (nbody, 1))
    for bi, b in enumerate(bodies):
        if bi == 0:
            continue
        if b["inertial"] is not None:
            it = b["inertial"]
            body_mass[bi], body_inertia[bi] = it["mass"], it["inertia"]
            body_ipos[bi], body_iquat[bi] = it["pos"], it["quat"]
            continue
        ms, cs, Is = [], [], []
        for gi in b["geoms"]:
            g = geoms[gi]
            if g["type"] == GEOM_PLANE:
                continue
            vol, Iu = _geom_volume_inertia(g["type"], g["size"])
            m = g["mass"] if g["mass"] is not None else g["density"] * vol
            if m <= 0 or vol <= 0:
                continue
            R = quat2mat(g["quat"])
            ms.append(m)
            cs.append(g["pos"])
            Is.append(R @ np.diag(Iu * (m / vol)) @ R.T)
        if not ms:
            continue
        mt = float(np.sum(ms))
        com = np.sum([m * c for m, c in zip(ms, cs)], axis=0) / mt
        It = np.zeros((3, 3))
        for m, c, Ig in zip(ms, cs, Is):
            dd = c - com
            It += Ig + m * (dd @ dd * np.eye(3) - np.outer(dd, dd))
        w, V = np.linalg.eigh(It)
        if np.allclose(It, np.diag(np.diag(It)), atol=1e-14 * max(1.0, np.abs(It).max())):
            w, V = np.diag(It).copy(), np.eye(3)
        if np.linalg.det(V) < 0:
            V[:, 2] = -V[:, 2]
        body_mass[bi], body_inertia[bi] = mt, w
        body_ipos[bi], body_iquat[bi] = com, mat2quat(V)
    A.update(body_mass=body_mass, body_inertia=body_inertia, body_ipos=body_ipos, body_iquat=body_iquat)

    # ---- actuators
    acts = []
    ae = root.find("actuator")
    if ae is not None:
        for c in ae:
            j = joints[names["joint"][c.get("joint")]]
            if j["type"] == JNT_FREE:
                raise ValueError("actuator on free joint unsupported")
            cr = _floats(c.get("ctrlrange"), 2, [0, 0])
            frng = _floats(c.get("forcerange"), 2, [0, 0])
            cl = c.get("ctrllimited")
            cl = (c.get("ctrlrange") is not None) if cl in (None, "auto") else cl == "true"
            fl = c.get("forcelimited")
            fl = (c.get("forcerange") is not None) if fl in (None, "auto") else fl == "true"
            if c.tag == "motor":
                gain, bias = 1.0, [0.0, 0.0, 0.0]
            elif c.tag == "position":
                kp = float(c.get("kp", 1))
                gain, bias = kp, [0.0, -kp, 0.0]
            else:
                raise ValueError(f"actuator type {c.tag} unsupported")
            acts.append(dict(name=c.get("name", f"act{len(acts)}"), dof=j["dofadr"], gain=gain, bias=bias,
                             ctrlrange=cr, forcerange=frng, ctrllimited=cl, forcelimited=fl,
                             gear=float(c.get("gear", "1").split()[0])))
            names["actuator"][acts[-1]["name"]] = len(acts) - 1
    nu = len(acts)
    A["act_dofid"] = np.array([a["dof"] for a in acts], I32)
    A["act_ctrllimited"] = np.array([int(a["ctrllimited"]) for a in acts], I32)
    A["act_forcelimited"] = np.array([int(a["forcelimited"]) for a in acts], I32)
    A["act_gain"] = np.array([a["gain"] for a in acts], float)
    A["act_bias"] = np.array([a["bias"] for a in acts], float).reshape(nu, 3)
    A["act_ctrlrange"] = np.array([a["ctrlrange"] for a in acts], float).reshape(nu, 2)
    A["act_forcerange"] = np.array([a["forcerange"] for a in acts], float).reshape(nu, 2)
    A["act_gear"] = np.array([a["gear"] for a in acts], float)

    # ---- candidate contact pairs (static filters; MuJoCo order: body pair, then geom ids)
    weld = np.zeros(nbody, I32)  # body_weldid: nearest ancestor-or-self with dofs, else world
    for bi in range(1, nbody):
        weld[bi] = bi if body_dofnum[bi] > 0 else weld[bodies[bi]["parent"]]
    excl = set()
    ce = root.find("contact")
    if ce is not None:
        for c in ce.findall("exclude"):
            b1, b2 = names["body"][c.get("body1")], names["body"][c.get("body2")]
            excl.add((min(b1, b2), max(b1, b2)))
    pairs = []
    for b1 in range(nbody):
        for b2 in range(b1 + 1, nbody):
            if (b1, b2) in excl:
                continue
            w1, w2 = weld[b1], weld[b2]
            if w1 == w2:
                continue
            if w1 and w2 and (weld[bodies[w1]["parent"]] == w2 or weld[bodies[w2]["parent"]] == w1):
                continue
            for g1 in bodies[b1]["geoms"]:
                for g2 in bodies[b2]["geoms"]:
                    ga, gb = geoms[g1], geoms[g2]
                    if not ((ga["contype"] & gb["conaffinity"]) or (gb["contype"] & ga["conaffinity"])):
                        continue
                    if ga["type"] > gb["type"]:  # MuJoCo orders the pair by geom type
                        pairs.append((g2, g1))
                    else:
                        pairs.append((g1, g2))
    npair = len(pairs)
    P = dict(pair_geom1=np.zeros(npair, I32), pair_geom2=np.zeros(npair, I32), pair_condim=np.zeros(npair, I32),
             pair_friction=np.zeros((npair, 5)), pair_solref=np.zeros((npair, 2)),
             pair_solimp=np.zeros((npair, 5)), pair_margin=np.zeros(npair), pair_gap=np.zeros(npair))
    for k, (g1, g2) in enumerate(pairs):
        ga, gb = geoms[g1], geoms[g2]
        P["pair_geom1"][k], P["pair_geom2"][k] = g1, g2
        if ga["priority"] != gb["priority"]:
            gp = ga if ga["priority"] > gb["priority"] else gb
            condim, fr, solref, solimp = gp["condim"], gp["friction"], gp["solref"], gp["solimp"]
        else:
            condim = max(ga["condim"], gb["condim"])
            fr = np.maximum(ga["friction"], gb["friction"])
            sm = ga["solmix"] + gb["solmix"]
            mix = 0.5 if sm < MINVAL else ga["solmix"] / sm
            if ga["solref"][0] > 0 and gb["solref"][0] > 0:
                solref = mix * ga["solref"] + (1 - mix) * gb["solref"]
            else:
                solref = np.minimum(ga["solref"], gb["solref"])
            solimp = mix * ga["solimp"] + (1 - mix) * gb["solimp"]
        P["pair_condim"][k] = condim
        P["pair_friction"][k] = [fr[0], fr[0], fr[1], fr[2], fr[2]]
        P["pair_solref"][k], P["pair_solimp"][k] = solref, solimp
        P["pair_margin"][k] = max(ga["margin"], gb["margin"])
        P["pair_gap"][k] = max(ga["gap"], gb["gap"])
    A.update(P)

    m = Model(opt=opt, names=names, arrays=A)
    m.opt.update(nq=nq, nv=nv, nu=nu, nbody=nbody, njnt=njnt, ngeom=ngeom, nsite=nsite, npair=npair, nM=int(nM))
    _set_const(m)
    return m


# ----------------------------------------------------------------------------- qpos0 constants
def kinematics(m: Model, qpos):
    """Plain numpy forward kinematics (used at compile time and by host-side helpers)."""
    nbody = m.nbody
    xpos, xquat = np.zeros((nbody, 3)), np.tile([1.0, 0, 0, 0], (nbody, 1))
    xanchor, xaxis = np.zeros((m.njnt, 3)), np.zeros((m.njnt, 3))
    for b in range(1, nbody):
        p = m.body_parentid[b]
        jn, ja = m.body_jntnum[b], m.body_jntadr[b]
        if jn == 1 and m.jnt_type[ja] == JNT_FREE:
            a = m.jnt_qposadr[ja]
            pos = np.array(qpos[a:a + 3], float)
            quat = np.array(qpos[a + 3:a + 7], float)
            quat /= np.linalg.norm(quat)
            xanchor[ja], xaxis[ja] = pos, [0, 0, 1]
        else:
            Rp = quat2mat(xquat[p])
            pos = xpos[p] + Rp @ m.body_pos[b]
            quat = quat_mul(xquat[p], m.body_quat[b])
            for k in range(jn):
                j = ja + k
                R = quat2mat(quat)
                axis = R @ m.jnt_axis[j]
                anchor = pos + R @ m.jnt_pos[j]
                dq = qpos[m.jnt_qposadr[j]] - m.qpos0[m.jnt_qposadr[j]]
                if m.jnt_type[j] == JNT_SLIDE:
                    pos = pos + axis * dq
                else:
                    quat = quat_mul(axisangle2quat(axis, dq), quat)  # world-axis rotation, pre-multiplied
                    pos = anchor - quat2mat(quat) @ m.jnt_pos[j]
                xanchor[j], xaxis[j] = anchor, axis
        xpos[b], xquat[b] = pos, quat / np.linalg.norm(quat)
    return xpos, xquat, xanchor, xaxis


def body_jacobian(m: Model, xpos, xquat, xanchor, xaxis, body, point):
    """6 x nv Jacobian (translational rows 0-2, rotational rows 3-5) of `point` fixed to `body`."""
    J = np.zeros((6, m.nv))
    b = body
    while b > 0:
        for k in range(m.body_jntnum[b]):
            j = m.body_jntadr[b] + k
            d = m.jnt_dofadr[j]
            t = m.jnt_type[j]
            if t == JNT_FREE:
                J[0:3, d:d + 3] = np.eye(3)
                R = quat2mat(xquat[b])
                for a in range(3):
                    ax = R[:, a]
                    J[3:6, d + 3 + a] = ax
                    J[0:3, d + 3 + a] = np.cross(ax, point - xpos[b])
            elif t == JNT_SLIDE:
                J[0:3, d] = xaxis[j]
            else:
                J[3:6, d] = xaxis[j]
                J[0:3, d] = np.cross(xaxis[j], point - xanchor[j])
        b = m.body_parentid[b]
    return J


def dense_mass_matrix(m: Model, qpos):
    """M = sum_b J_b^T diag(m, I_b) J_b  (+ armature).  Deliberately not CRB: independent check."""
    xpos, xquat, xanchor, xaxis = kinematics(m, qpos)
    M = np.diag(np.array(m.dof_armature, float))
    for b in range(1, m.nbody):
        if m.body_mass[b] <= 0:
            continue
        R = quat2mat(xquat[b])
        com = xpos[b] + R @ m.body_ipos[b]
        Ri = R @ quat2mat(m.body_iquat[b])
        Iw = Ri @ np.diag(m.body_inertia[b]) @ Ri.T
        J = body_jacobian(m, xpos, xquat, xanchor, xaxis, b, com)
        M += m.body_mass[b] * J[:3].T @ J[:3] + J[3:].T @ Iw @ J[3:]
    return M


def _set_const(m: Model):
    nv = m.nv
    A = m.arrays
    if nv == 0:
        A["dof_invweight0"], A["body_invweight0"] = np.zeros(0), np.zeros((m.nbody, 2))
        m.opt["meaninertia"] = 1.0
        return
    M = dense_mass_matrix(m, m.qpos0)
    Minv = np.linalg.inv(M)
    m.opt["meaninertia"] = float(np.trace(M) / nv)
    dw = np.zeros(nv)
    for j in range(m.njnt):
        d = m.jnt_dofadr[j]
        if m.jnt_type[j] == JNT_FREE:
            dw[d:d + 3] = np.mean(np.diag(Minv)[d:d + 3])
            dw[d + 3:d + 6] = np.mean(np.diag(Minv)[d + 3:d + 6])
        else:
            dw[d] = Minv[d, d]
    A["dof_invweight0"] = dw
    xpos, xquat, xanchor, xaxis = kinematics(m, m.qpos0)
    bw = np.zeros((m.nbody, 2))
    for b in range(1, m.nbody):
        R = quat2mat(xquat[b])
        com = xpos[b] + R @ m.body_ipos[b]
        J = body_jacobian(m, xpos, xquat, xanchor, xaxis, b, com)
        if not np.any(J):
            continue
        Ainv = J @ Minv @ J.T
        bw[b, 0] = max(MINVAL, np.trace(Ainv[:3, :3]) / 3)
        bw[b, 1] = max(MINVAL, np.trace(Ainv[3:, 3:]) / 3)
    A["body_invweight0"] = bw
