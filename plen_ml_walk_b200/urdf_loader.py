"""PLEN URDF + foot STL -> structure-of-arrays articulation tables for the sm_100a kernels.

Host-side loader of the hot path (reference: `p.loadURDF("plen.urdf", ...)`,
plen_bullet/src/plen_bullet/plen_env.py:312-315; model plen_bullet/src/plen.urdf:504-1489; foot collision
meshes plen_ros/meshes_bin/{r,l}foot.stl, plen.urdf:1097 and :1263).

Design (B200-first, not Bullet's): the 14 fixed-joint links are folded exactly into their parents as composite
rigid bodies, leaving 19 moving bodies / 24 generalized velocities, laid out so that ONE WARP owns one robot:

    lane 0..5   the six base velocity coordinates (omega_world xyz, v_world xyz); lane 0 also carries the torso body
    lane 6..11  right leg   (r_hip, r_thigh, r_knee, r_shin, r_ankle, r_foot)
    lane 12..17 left leg
    lane 18..20 right arm   (r_shoulder, rs_servo, r_elbow)
    lane 21..23 left arm

so lane l >= 6 is simultaneously "body l-5", "joint l-6" (action/observation order, plen_env.py:318-320) and
"generalized velocity l".  Each limb is a serial chain, which lets the kernels replace tree traversals by segmented
warp scans.

Inertia follows what PyBullet does without URDF_USE_INERTIA_FROM_FILE: the URDF <inertia>
is ignored and each link gets the box inertia of its collision AABB grown by the collision margins.

The packaged tables in `data/plen_model.json` were produced by running this module on the read-only reference
checkout (`python -m plen_ml_walk_b200.urdf_loader /root/reference`); the GPU box has no reference tree.
"""
from __future__ import annotations

import json
import os
import struct
import xml.etree.ElementTree as ET
from dataclasses import dataclass, field

import numpy as np

N_LANES = 24
N_JOINTS = 18
CHAIN_STARTS = (6, 12, 18, 21)
CHAIN_LENGTHS = (6, 6, 3, 3)

# joint order of the action / observation vectors (plen_env.py:547-554)
JOINT_NAMES = (
    "rb_servo_r_hip", "r_hip_r_thigh", "r_thigh_r_knee", "r_knee_r_shin", "r_shin_r_ankle", "r_ankle_r_foot",
    "lb_servo_l_hip", "l_hip_l_thigh", "l_thigh_l_knee", "l_knee_l_shin", "l_shin_l_ankle", "l_ankle_l_foot",
    "torso_r_shoulder", "r_shoulder_rs_servo", "re_servo_r_elbow",
    "torso_l_shoulder", "l_shoulder_ls_servo", "le_servo_l_elbow",
)
FOOT_JOINTS = ("r_ankle_r_foot", "l_ankle_l_foot")          # Bullet links 11 and 19 (plen_env.py:446-467)

SHAPE_MARGIN = 1.0e-3        # collision margin of every URDF shape; convex hulls are inflated by it
COMPOUND_MARGIN = 1.0e-3     # margin of the per-link compound shape, added once more to its AABB
BREAKING_FACTOR = 0.02       # contact breaking threshold = factor * (|aabb centre| + aabb radius)
MAX_BOXES = 32               # box colliders besides the feet (plen.urdf: torso + 30 links), one per lane of the warp

_DATA_JSON = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "plen_model.json")


def _rot_rpy(r, p, y):
    cr, sr, cp, sp, cy, sy = np.cos(r), np.sin(r), np.cos(p), np.sin(p), np.cos(y), np.sin(y)
    rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]])
    ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
    rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]])
    return rz @ ry @ rx


def _floats(text):
    return np.array([float(t) for t in text.split()], dtype=np.float64)


def _origin(elem):
    o = elem.find("origin") if elem is not None else None
    if o is None:
        return np.eye(3), np.zeros(3)
    return _rot_rpy(*_floats(o.get("rpy", "0 0 0"))), _floats(o.get("xyz", "0 0 0"))


def _stl_points(path):
    raw = open(path, "rb").read()
    n = struct.unpack_from("<I", raw, 80)[0]
    if len(raw) != 84 + 50 * n:
        raise ValueError("%s is not a binary STL" % path)
    rec = np.frombuffer(raw, dtype=np.uint8, count=50 * n, offset=84).reshape(n, 50)
    pts = rec[:, 12:48].copy().view("<f4").reshape(-1, 3).astype(np.float64)
    return np.unique(pts, axis=0)


def _hull_vertices(pts):
    """The points of a collider's vertex cloud that are vertices of its convex hull, in the order of the (sorted) cloud --
    what the support-vertex search of the persistent sole manifold walks (plen_config.sole_manifold)."""
    from scipy.spatial import ConvexHull
    return pts[np.sort(ConvexHull(pts).vertices)]


def _largest_quad(poly):
    """Indices of the 4 polygon vertices (ccw order) enclosing the largest area -- O(n^4) is fine for n = 32."""
    n = len(poly)
    x, y = poly[:, 0], poly[:, 1]
    best, pick = -1.0, None
    for a in range(n):
        for b in range(a + 1, n):
            for c in range(b + 1, n):
                for d in range(c + 1, n):
                    i = (a, b, c, d)
                    s = 0.0
                    for k in range(4):
                        s += x[i[k]] * y[i[(k + 1) % 4]] - x[i[(k + 1) % 4]] * y[i[k]]
                    s = abs(s) * 0.5
                    if s > best + 1e-15:
                        best, pick = s, i
    return list(pick)


@dataclass
class _Link:
    name: str
    mass: float
    com: np.ndarray
    inertia_diag: np.ndarray
    aabb_disc: float
    hull: np.ndarray | None
    hull_margin: float = 0.0     # collision margin the contact surface is inflated by (convex hulls: SHAPE_MARGIN, boxes: 0)
    boxes: list = field(default_factory=list)     # (rotation, centre, half extents) of every box collider, link frame


@dataclass
class PlenModel:
    """float64 master copy of the per-lane tables; `.as_c()` packs float32 for the C ABI (include/plen_b200.h)."""
    R_pj: np.ndarray = field(default_factory=lambda: np.tile(np.eye(3), (N_LANES, 1, 1)))
    p_pj: np.ndarray = field(default_factory=lambda: np.zeros((N_LANES, 3)))
    axis: np.ndarray = field(default_factory=lambda: np.zeros((N_LANES, 3)))
    mass: np.ndarray = field(default_factory=lambda: np.zeros(N_LANES))
    com: np.ndarray = field(default_factory=lambda: np.zeros((N_LANES, 3)))
    inertia: np.ndarray = field(default_factory=lambda: np.zeros((N_LANES, 6)))   # xx yy zz xy xz yz about the com
    lower: np.ndarray = field(default_factory=lambda: np.zeros(N_LANES))
    upper: np.ndarray = field(default_factory=lambda: np.zeros(N_LANES))
    chain_start: np.ndarray = field(default_factory=lambda: np.zeros(N_LANES, dtype=np.int32))
    foot_lane: np.ndarray = field(default_factory=lambda: np.zeros(2, dtype=np.int32))
    foot_pts: np.ndarray = field(default_factory=lambda: np.zeros((2, 4, 3)))
    foot_break: np.ndarray = field(default_factory=lambda: np.zeros(2))
    foot_margin: float = 0.001     # collision margin of the foot shape (convex hull: 1 mm, box: 0) -> plen_config.hull_margin
    foot_hull: list = field(default_factory=lambda: [[], []])   # convex-hull vertices of either foot collider, foot body frame
    # box colliders of every link except the feet (ground contact of knees / hands / torso ...), in Bullet's
    # link order (base first, then DFS pre-order over the joints in file order); pose in the frame of the BODY (lane) the
    # link is folded into
    n_boxes: int = 0
    box_lane: np.ndarray = field(default_factory=lambda: np.zeros(MAX_BOXES, dtype=np.int32))
    box_center: np.ndarray = field(default_factory=lambda: np.zeros((MAX_BOXES, 3)))
    box_rot: np.ndarray = field(default_factory=lambda: np.tile(np.eye(3), (MAX_BOXES, 1, 1)))
    box_half: np.ndarray = field(default_factory=lambda: np.zeros((MAX_BOXES, 3)))
    box_rest: np.ndarray = field(default_factory=lambda: np.zeros(MAX_BOXES))   # 0: the base link (restitution 0), 1: any other
    box_names: list = field(default_factory=list)
    body_links: list = field(default_factory=list)     # names of the URDF links folded into each lane's body
    total_mass: float = 0.0

    def to_json(self, path=_DATA_JSON):
        d = {k: (v.tolist() if isinstance(v, np.ndarray) else v) for k, v in self.__dict__.items()}
        os.makedirs(os.path.dirname(path), exist_ok=True)
        with open(path, "w") as f:
            json.dump(d, f, indent=1)

    @staticmethod
    def from_json(path=_DATA_JSON):
        with open(path) as f:
            d = json.load(f)
        m = PlenModel()
        for k, v in d.items():
            cur = getattr(m, k)
            setattr(m, k, np.array(v, dtype=cur.dtype) if isinstance(cur, np.ndarray) else v)
        return m


def _read_links(root, mesh_dir):
    out = {}
    for le in root.findall("link"):
        ine = le.find("inertial")
        mass = float(ine.find("mass").get("value"))
        iR, com = _origin(ine)
        if not np.allclose(iR, np.eye(3)):
            raise NotImplementedError("rotated inertial frames are not used by plen.urdf")
        lo, hi, hull, margin, boxes = None, None, None, 0.0, []
        for col in le.findall("collision"):
            cR, cx = _origin(col)
            g = col.find("geometry")[0]
            if g.tag == "box":
                half = _floats(g.get("size")) / 2
                boxes.append((cR, cx, half))
                ext = np.abs(cR) @ half
                a, b = cx - ext, cx + ext
                if hull is None:      # box feet (plen_new.urdf): the 8 corners stand in for the hull; Bullet boxes keep their size
                    corners = np.array([[sx, sy, sz] for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)]) * half
                    hull, margin = corners @ cR.T + cx, 0.0
            elif g.tag == "mesh":
                pts = _stl_points(os.path.join(mesh_dir, os.path.basename(g.get("filename"))))
                hull, margin = (pts * _floats(g.get("scale", "1 1 1"))) @ cR.T + cx, SHAPE_MARGIN
                a, b = hull.min(0) - SHAPE_MARGIN, hull.max(0) + SHAPE_MARGIN
            else:
                raise NotImplementedError(g.tag)
            lo = a if lo is None else np.minimum(lo, a)
            hi = b if hi is None else np.maximum(hi, b)
        lo, hi = lo - COMPOUND_MARGIN, hi + COMPOUND_MARGIN
        e = hi - lo
        diag = mass / 12.0 * np.array([e[1] ** 2 + e[2] ** 2, e[0] ** 2 + e[2] ** 2, e[0] ** 2 + e[1] ** 2])
        disc = float(np.linalg.norm((lo + hi) / 2 - com) + np.linalg.norm(e) / 2)
        out[le.get("name")] = _Link(le.get("name"), mass, com, diag, disc, hull, margin, boxes)
    return out


def load_plen_model(urdf_path, mesh_dir) -> PlenModel:
    root = ET.parse(urdf_path).getroot()
    links = _read_links(root, mesh_dir)
    joints = {j.get("name"): j for j in root.findall("joint")}
    by_parent, child_of = {}, {}
    for j in joints.values():
        by_parent.setdefault(j.find("parent").get("link"), []).append(j)
        child_of[j.find("child").get("link")] = j

    model = PlenModel()
    model.total_mass = float(sum(l.mass for l in links.values()))
    lane_of_link = {}     # root link of each moving body -> lane

    def fold(root_link):
        """All links rigidly attached to root_link, with their pose in root_link's frame."""
        acc = [(root_link, np.eye(3), np.zeros(3))]
        k = 0
        while k < len(acc):
            name, R, p = acc[k]
            k += 1
            for j in by_parent.get(name, []):
                if j.get("type") == "fixed":
                    jR, jp = _origin(j)
                    acc.append((j.find("child").get("link"), R @ jR, p + R @ jp))
        return acc

    def set_body(lane, root_link):
        parts = fold(root_link)
        m = sum(links[n].mass for n, _, _ in parts)
        c = sum(links[n].mass * (p + R @ links[n].com) for n, R, p in parts) / m
        I = np.zeros((3, 3))
        for n, R, p in parts:
            d = p + R @ links[n].com - c
            I += R @ np.diag(links[n].inertia_diag) @ R.T + links[n].mass * (d @ d * np.eye(3) - np.outer(d, d))
        model.mass[lane], model.com[lane] = m, c
        model.inertia[lane] = [I[0, 0], I[1, 1], I[2, 2], I[0, 1], I[0, 2], I[1, 2]]
        model.body_links.append([n for n, _, _ in parts])
        return {n: (R, p) for n, R, p in parts}

    base = [n for n in links if n not in child_of]
    assert base == ["torso"], base
    model.body_links = []
    poses = {0: set_body(0, "torso")}
    lane_of_link["torso"] = 0
    body_of_link = {n: 0 for n in poses[0]}
    order = [None] * N_LANES
    for ji, jn in enumerate(JOINT_NAMES):
        order[6 + ji] = jn
    # lanes must be filled parents-first; JOINT_NAMES already lists every chain root-to-tip
    model.body_links = [model.body_links[0]] + [[] for _ in range(N_LANES - 1)]
    for lane in range(6, N_LANES):
        j = joints[order[lane]]
        assert j.get("type") == "revolute", order[lane]
        parent_link, child_link = j.find("parent").get("link"), j.find("child").get("link")
        pbody = body_of_link[parent_link]
        cs = [s for s in CHAIN_STARTS if s <= lane][-1]
        assert pbody == (0 if lane == cs else lane - 1), (order[lane], pbody)   # serial chains
        R_fix, p_fix = poses[pbody][parent_link]
        jR, jp = _origin(j)
        model.R_pj[lane], model.p_pj[lane] = R_fix @ jR, p_fix + R_fix @ jp
        model.axis[lane] = _floats(j.find("axis").get("xyz"))
        model.axis[lane] /= np.linalg.norm(model.axis[lane])
        lim = j.find("limit")
        model.lower[lane], model.upper[lane] = float(lim.get("lower")), float(lim.get("upper"))
        model.chain_start[lane] = cs
        saved = model.body_links
        model.body_links = []
        poses[lane] = set_body(lane, child_link)
        saved[lane] = model.body_links[0]
        model.body_links = saved
        for n in poses[lane]:
            body_of_link[n] = lane
    assert len(body_of_link) == len(links), "unreached links"

    # ---- box colliders of the non-foot links, Bullet link order (base, then DFS pre-order, children in joint-file order)
    foot_links = {joints[jn].find("child").get("link") for jn in FOOT_JOINTS}
    dfs = []

    def walk(name):
        dfs.append(name)
        for j in by_parent.get(name, []):
            walk(j.find("child").get("link"))

    walk("torso")
    for name in dfs:
        if name in foot_links:
            continue
        lane = body_of_link[name]
        R, p = poses[lane][name]
        for cR, cx, half in links[name].boxes:
            b = model.n_boxes
            if b >= MAX_BOXES:
                raise ValueError("more than %d box colliders" % MAX_BOXES)
            model.box_lane[b], model.box_center[b], model.box_rot[b], model.box_half[b] = lane, p + R @ cx, R @ cR, half
            model.box_rest[b] = 0.0 if name == "torso" else 1.0
            model.box_names.append(name)
            model.n_boxes += 1

    for f, jn in enumerate(FOOT_JOINTS):
        lane = 6 + JOINT_NAMES.index(jn)
        link = links[joints[jn].find("child").get("link")]
        hull = link.hull
        zmin = hull[:, 2].min()
        sole = hull[np.abs(hull[:, 2] - zmin) < 1e-6]
        ctr = sole[:, :2].mean(0)
        sole = sole[np.argsort(np.arctan2(sole[:, 1] - ctr[1], sole[:, 0] - ctr[0]))]
        quad = sole[_largest_quad(sole)]
        ang = np.arctan2(quad[:, 1] - ctr[1], quad[:, 0] - ctr[0])
        quad = np.roll(quad, -int(np.argmin(np.where(ang >= 0, ang, ang + 2 * np.pi))), axis=0)
        model.foot_lane[f] = lane
        model.foot_pts[f] = quad
        model.foot_break[f] = BREAKING_FACTOR * link.aabb_disc
        model.foot_margin = float(link.hull_margin)     # -> plen_config.hull_margin (PlenVecEnv applies it)
        model.foot_hull[f] = _hull_vertices(hull).tolist()
    return model


def packaged_model(name: str = "plen") -> PlenModel:
    """The tables shipped with the package (generated from the reference URDFs by this module's __main__):
    "plen" = plen_bullet/src/plen.urdf (the one plen_env.py:312 loads; convex-hull feet from the STL meshes),
    "plen_new" = plen_bullet/src/plen_new.urdf (box feet, +-1.0 rad joint limits; unused by the reference's env)."""
    if name == "plen":
        return PlenModel.from_json(_DATA_JSON)
    if name == "plen_new":
        return PlenModel.from_json(_DATA_JSON.replace("plen_model.json", "plen_new_model.json"))
    raise ValueError("unknown packaged model %r" % name)


if __name__ == "__main__":
    import sys
    ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    mdl = load_plen_model(os.path.join(ref, "plen_bullet/src/plen.urdf"), os.path.join(ref, "plen_ros/meshes_bin"))
    mdl.to_json()
    load_plen_model(os.path.join(ref, "plen_bullet/src/plen_new.urdf"), os.path.join(ref, "plen_ros/meshes_bin")).to_json(
        _DATA_JSON.replace("plen_model.json", "plen_new_model.json"))
    print("total mass %.6f kg, bodies:" % mdl.total_mass)
    for lane in [0] + list(range(6, N_LANES)):
        print(lane, "%.6f" % mdl.mass[lane], mdl.body_links[lane])
