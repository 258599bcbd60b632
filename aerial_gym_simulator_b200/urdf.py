"""Minimal URDF reader: inertials, fixed-joint tree, box visuals.

Replaces what the reference obtains from Isaac Gym asset properties
(robots/robot_manager.py:295-435: per-body mass / COM / inertia, body poses) and from
urdfpy + trimesh (assets/warp_asset.py:19-60: visual meshes with link transforms).
Only what the hot paths consume is parsed; there is no physics here."""
from __future__ import annotations

import math
import os
import xml.etree.ElementTree as ET
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np


def rpy_to_matrix(rpy) -> np.ndarray:
    """URDF fixed-axis roll-pitch-yaw: R = Rz(yaw) Ry(pitch) Rx(roll)."""
    r, p, y = (float(v) for v in rpy)
    cr, sr, cp, sp, cy, sy = math.cos(r), math.sin(r), math.cos(p), math.sin(p), math.cos(y), math.sin(y)
    return np.array([
        [cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr],
        [sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr],
        [-sp, cp * sr, cp * cr],
    ])


def _vec(s: Optional[str], n=3, default=0.0):
    if s is None:
        return np.full(n, default)
    return np.array([float(x) for x in s.split()], dtype=np.float64)


@dataclass
class Visual:
    kind: str  # "box" | "sphere" | "cylinder" | "mesh"
    size: np.ndarray
    R: np.ndarray
    p: np.ndarray
    filename: Optional[str] = None


@dataclass
class Link:
    name: str
    mass: float = 0.0
    com: np.ndarray = field(default_factory=lambda: np.zeros(3))
    com_R: np.ndarray = field(default_factory=lambda: np.eye(3))
    inertia: np.ndarray = field(default_factory=lambda: np.zeros((3, 3)))
    visuals: List[Visual] = field(default_factory=list)
    collisions: List[Visual] = field(default_factory=list)


@dataclass
class Joint:
    name: str
    parent: str
    child: str
    R: np.ndarray
    p: np.ndarray
    kind: str = "fixed"


@dataclass
class UrdfModel:
    name: str
    links: Dict[str, Link]
    joints: List[Joint]
    root: str

    def body_order(self) -> List[str]:
        """Rigid-body index order used by the reference configs' application_mask
        (e.g. base_quad_config.py:163: bodies 1-4 arms, 5-8 motors): root first, then the
        children of each link in name order, depth first."""
        children: Dict[str, List[str]] = {}
        for j in self.joints:
            children.setdefault(j.parent, []).append(j.child)
        out: List[str] = []

        def visit(n):
            out.append(n)
            for c in sorted(children.get(n, [])):
                visit(c)

        visit(self.root)
        return out

    def link_transforms(self) -> Dict[str, Tuple[np.ndarray, np.ndarray]]:
        """(R, p) of every link frame in the root frame, joints at their zero position."""
        tf = {self.root: (np.eye(3), np.zeros(3))}
        pending = list(self.joints)
        while pending:
            progressed = False
            for j in list(pending):
                if j.parent in tf:
                    Rp, pp = tf[j.parent]
                    tf[j.child] = (Rp @ j.R, pp + Rp @ j.p)
                    pending.remove(j)
                    progressed = True
            if not progressed:
                raise ValueError("URDF joint graph is not a tree rooted at " + self.root)
        return tf

    def composite_inertia(self) -> Tuple[float, np.ndarray, np.ndarray]:
        """Total mass, COM (root frame) and inertia about the COM in root-frame axes: the
        parallel-axis accumulation of robots/robot_manager.py:313-426."""
        tf = self.link_transforms()
        m_tot, c_sum = 0.0, np.zeros(3)
        for l in self.links.values():
            R, p = tf[l.name]
            c_sum += l.mass * (p + R @ l.com)
            m_tot += l.mass
        com = c_sum / m_tot if m_tot > 0 else np.zeros(3)
        J = np.zeros((3, 3))
        for l in self.links.values():
            R, p = tf[l.name]
            Rl = R @ l.com_R
            d = (p + R @ l.com) - com
            J += Rl @ l.inertia @ Rl.T + l.mass * (float(d @ d) * np.eye(3) - np.outer(d, d))
        return m_tot, com, J


def _parse_geometry(el, base_dir) -> Optional[Visual]:
    geom = el.find("geometry")
    if geom is None:
        return None
    org = el.find("origin")
    R = rpy_to_matrix(_vec(org.get("rpy") if org is not None else None))
    p = _vec(org.get("xyz") if org is not None else None)
    b = geom.find("box")
    if b is not None:
        return Visual("box", _vec(b.get("size")), R, p)
    s = geom.find("sphere")
    if s is not None:
        return Visual("sphere", np.array([float(s.get("radius"))]), R, p)
    c = geom.find("cylinder")
    if c is not None:
        return Visual("cylinder", np.array([float(c.get("radius")), float(c.get("length"))]), R, p)
    m = geom.find("mesh")
    if m is not None:
        return Visual("mesh", _vec(m.get("scale"), default=1.0), R, p, os.path.join(base_dir, m.get("filename", "")))
    return None


def parse_urdf(path: str) -> UrdfModel:
    root_el = ET.parse(path).getroot()
    base_dir = os.path.dirname(os.path.abspath(path))
    links: Dict[str, Link] = {}
    for le in root_el.findall("link"):
        l = Link(le.get("name"))
        ine = le.find("inertial")
        if ine is not None:
            m = ine.find("mass")
            l.mass = float(m.get("value")) if m is not None else 0.0
            org = ine.find("origin")
            if org is not None:
                l.com = _vec(org.get("xyz"))
                l.com_R = rpy_to_matrix(_vec(org.get("rpy")))
            ie = ine.find("inertia")
            if ie is not None:
                g = lambda k: float(ie.get(k, 0.0))
                l.inertia = np.array([[g("ixx"), g("ixy"), g("ixz")], [g("ixy"), g("iyy"), g("iyz")], [g("ixz"), g("iyz"), g("izz")]])
        for ve in le.findall("visual"):
            v = _parse_geometry(ve, base_dir)
            if v is not None:
                l.visuals.append(v)
        for ce in le.findall("collision"):
            v = _parse_geometry(ce, base_dir)
            if v is not None:
                l.collisions.append(v)
        links[l.name] = l
    joints: List[Joint] = []
    children = set()
    for je in root_el.findall("joint"):
        org = je.find("origin")
        R = rpy_to_matrix(_vec(org.get("rpy") if org is not None else None))
        p = _vec(org.get("xyz") if org is not None else None)
        j = Joint(je.get("name"), je.find("parent").get("link"), je.find("child").get("link"), R, p, je.get("type", "fixed"))
        joints.append(j)
        children.add(j.child)
    roots = [n for n in links if n not in children]
    if len(roots) != 1:
        raise ValueError(f"{path}: expected exactly one root link, found {roots}")
    return UrdfModel(root_el.get("name", ""), links, joints, roots[0])


def box_visual_triangles(model: UrdfModel, use_collision=False):
    """Per-link list of [12,9] root-frame triangle arrays for every <box> visual (the obstacle
    assets the reference ships are all boxes).  Returns [(link_name, triangles, obb16)]."""
    from .hp2 import box_obb, box_triangles

    tf = model.link_transforms()
    out = []
    for name in model.body_order():
        l = model.links[name]
        R, p = tf[name]
        for v in (l.collisions if use_collision else l.visuals):
            if v.kind != "box":
                continue
            t = box_triangles(v.size).reshape(-1, 3).astype(np.float64)
            t = (R @ (v.R @ t.T + v.p[:, None]) + p[:, None]).T
            out.append((name, t.reshape(-1, 9).astype(np.float32), box_obb(v.size, R @ v.R, R @ v.p + p)))
    return out


def visual_parts(model: UrdfModel, use_collision=False, max_tris=12):
    """Every <box> / <cylinder> visual of the model as ray-cast scene parts: [(link_name, link_index, triangles [n<=max_tris,9] in the
    root frame, obb16)].  A box is one part (its 12 triangles + the oriented box the ray-caster uses for the entry-face shortcut,
    kind 2); a cylinder is tessellated (hp2.cylinder_triangles) and cut into parts of <= max_tris triangles that share the cylinder's
    bounding box (kind 1: culling only).  link_index counts the links that carry a visual, in urdfpy's link order -- the per-link
    segmentation counter of assets/warp_asset.py:44-70.  Spheres / meshes are not carried (no shipped obstacle uses them)."""
    from .hp2 import box_obb, box_triangles, cylinder_triangles

    tf = model.link_transforms()
    out, link_index = [], 0
    for name in model.body_order():
        l = model.links[name]
        R, p = tf[name]
        had_visual = False
        for v in (l.collisions if use_collision else l.visuals):
            if v.kind == "box":
                t, obb = box_triangles(v.size), box_obb(v.size, R @ v.R, R @ v.p + p)
            elif v.kind == "cylinder":
                t = cylinder_triangles(v.size[0], v.size[1])
                obb = box_obb((2 * v.size[0], 2 * v.size[0], v.size[1]), R @ v.R, R @ v.p + p)
                obb[15] = 1.0
            else:
                continue
            had_visual = True
            t = t.reshape(-1, 3).astype(np.float64)
            t = (R @ (v.R @ t.T + v.p[:, None]) + p[:, None]).T.reshape(-1, 9).astype(np.float32)
            for i in range(0, len(t), max_tris):
                out.append((name, link_index, t[i:i + max_tris], obb))
        link_index += int(had_visual)
    return out
