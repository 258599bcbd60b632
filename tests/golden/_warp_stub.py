"""A ~150-line stand-in for `warp` (warp-lang 1.0.0, not installable here) that lets the reference's OWN sensor classes and kernels run
on CPU, unmodified: `WarpCam`, `WarpLidar`, `WarpStereoCam`, `WarpNormalFaceID*` (sensors/warp/*.py) and every `@wp.kernel` of
sensors/warp/warp_kernels/*.py.  TEST INFRASTRUCTURE (this container only, used by make_golden_hp2.py to generate fixtures).

What is the reference's: ray generation, frame conventions, depth / range multiplier, far-plane handling, miss values, segmentation
lookup (first vertex of the hit face, x component of the mesh "velocity"), point-cloud frames, normal frames, the stereo logic,
the camera matrices and the LiDAR ray table.  What is NOT: `wp.mesh_query_ray` itself -- Warp's BVH traversal is replaced by a
brute-force closest hit over all triangles (float64 Moeller-Trumbore, first index wins ties); its returned normal is the unit
geometric normal cross(v1 - v0, v2 - v0) and `sign` = +1 for a front-face hit, as in Warp's documentation.

A kernel is plain Python here: `wp.launch` calls it once per index of `dim` with `wp.tid()` returning that index; vectors are float32
numpy arrays, so the arithmetic is fp32 like the device code (not bit-identical to a GPU: no FMA contraction is modelled).  Warp
passes the query's results through reference arguments, Python cannot: the `if wp.mesh_query_ray(mesh, o, d, max_t, t, u, v, sign, n,
f):` statements are rewritten (ast) into a tuple assignment followed by the same `if`."""
import ast
import inspect
import sys
import textwrap
import types

import numpy as np

f32 = np.float32
_tid = ()


def _v(*a):
    return np.array(a, dtype=np.float32)


class _Stub(types.ModuleType):
    pass


wp = _Stub("warp")
wp.float32, wp.uint64 = np.float32, np.uint64
wp.int32 = lambda x=0: np.int32(int(x))
wp.constant = lambda x: x
wp.array = wp.array2d = lambda *a, **k: (a[0] if a else None)  # an annotation (keywords only) or wp.array(list_of_meshes, dtype=...)
wp.vec3 = lambda *a: _v(*a) if a else np.zeros(3, np.float32)
wp.quat = lambda *a: _v(*a) if a else np.zeros(4, np.float32)
wp.mat44 = lambda *a: np.array(a, dtype=np.float32).reshape(4, 4)
wp.inverse = lambda m: np.linalg.inv(m.astype(np.float64)).astype(np.float32)
wp.tid = lambda: _tid
wp.dot = lambda a, b: f32(f32(f32(a[0] * b[0]) + f32(a[1] * b[1])) + f32(a[2] * b[2]))
wp.length = lambda a: f32(np.sqrt(wp.dot(a, a)))
wp.normalize = lambda a: (a / wp.length(a)).astype(np.float32)
wp.cross = lambda a, b: _v(a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0])
wp.transform_vector = lambda m, v: _v(*[f32(f32(f32(m[i, 0] * v[0]) + f32(m[i, 1] * v[1])) + f32(m[i, 2] * v[2])) for i in range(3)])
wp.quat_inverse = lambda q: _v(-q[0], -q[1], -q[2], q[3])
wp.mesh_get = lambda mesh: mesh
wp.capture_begin = wp.capture_end = wp.capture_launch = lambda *a, **k: None


def _quat_rotate(q, x):  # warp/native/quat.h: x (2 w^2 - 1) + cross(q.xyz, x) w 2 + q.xyz dot(q.xyz, x) 2
    qv, w = q[:3], q[3]
    return (x * f32(f32(2.0) * w * w - f32(1.0)) + wp.cross(qv, x) * w * f32(2.0) + qv * wp.dot(qv, x) * f32(2.0)).astype(np.float32)


wp.quat_rotate = _quat_rotate


class ScopedTimer:
    def __init__(self, *a, **k): pass
    def __enter__(self): return self
    def __exit__(self, *a): return False


wp.ScopedTimer = ScopedTimer


def _from_torch(t, dtype=None):
    return None if t is None else t.numpy()  # shares memory with the (CPU) torch tensor: the kernels write the caller's buffers


wp.from_torch = _from_torch
wp.to_torch = lambda a: __import__("torch").from_numpy(a)


class Mesh:
    """points [V,3] f32, indices [3F] i32, velocities [V,3] f32 (segmentation id in x: warp_env_manager.py:76-80)"""

    def __init__(self, points, indices, velocities):
        self.points, self.indices, self.velocities = np.asarray(points, np.float32), np.asarray(indices, np.int32), np.asarray(velocities, np.float32)
        p = self.points.astype(np.float64)[self.indices.reshape(-1, 3)]
        self._v0, self._e1, self._e2 = p[:, 0], p[:, 1] - p[:, 0], p[:, 2] - p[:, 0]
        n = np.cross(self._e1, self._e2)
        self._n = n / np.maximum(np.linalg.norm(n, axis=1, keepdims=True), 1e-300)


def _mesh_query_ray(mesh, ro, rd, max_t):
    o, d = ro.astype(np.float64), rd.astype(np.float64)
    p = np.cross(d, mesh._e2)
    det = (mesh._e1 * p).sum(1)
    ok = np.abs(det) > 1e-12
    inv = np.where(ok, 1.0 / np.where(ok, det, 1.0), 0.0)
    s = o - mesh._v0
    u = (s * p).sum(1) * inv
    q = np.cross(s, mesh._e1)
    v = (q * d).sum(1) * inv
    t = (q * mesh._e2).sum(1) * inv
    hit = ok & (u >= 0) & (v >= 0) & (u + v <= 1) & (t >= 0) & (t < float(max_t))
    if not hit.any():
        return False, f32(0), f32(0), f32(0), f32(0), np.zeros(3, np.float32), 0
    tt = np.where(hit, t, np.inf)
    f = int(np.argmin(tt))
    sign = f32(1.0) if det[f] > 0 else f32(-1.0)
    return True, f32(t[f]), f32(u[f]), f32(v[f]), sign, mesh._n[f].astype(np.float32), f


wp._mesh_query_ray = _mesh_query_ray


class _Rewrite(ast.NodeTransformer):
    """`if [not] wp.mesh_query_ray(mesh, o, d, max_t, t, u, v, sign, n, f):` ->
           _qN = wp._mesh_query_ray(mesh, o, d, max_t)
           if _qN[0]: t, u, v, sign, n, f = _qN[1:]        (a miss leaves the reference arguments untouched, as in Warp)
           if [not] _qN[0]: ..."""
    count = 0

    def visit_If(self, node):
        self.generic_visit(node)
        c, neg = node.test, False
        if isinstance(c, ast.UnaryOp) and isinstance(c.op, ast.Not):
            c, neg = c.operand, True
        if isinstance(c, ast.Call) and isinstance(c.func, ast.Attribute) and c.func.attr == "mesh_query_ray":
            _Rewrite.count += 1
            q = f"_q{_Rewrite.count}"
            call = ast.Call(func=ast.Attribute(value=ast.Name("wp", ast.Load()), attr="_mesh_query_ray", ctx=ast.Load()), args=c.args[:4], keywords=[])
            hit = ast.Subscript(ast.Name(q, ast.Load()), ast.Constant(0), ast.Load())
            outs = ast.Tuple([ast.Name(a.id, ast.Store()) for a in c.args[4:10]], ast.Store())
            take = ast.If(test=hit, body=[ast.Assign([outs], ast.Subscript(ast.Name(q, ast.Load()), ast.Slice(ast.Constant(1), None, None), ast.Load()))], orelse=[])
            node.test = ast.UnaryOp(ast.Not(), hit) if neg else hit
            return [ast.Assign([ast.Name(q, ast.Store())], call), take, node]
        return node


    def visit_Expr(self, node):  # a bare `wp.mesh_query_ray(...)` statement (normal / faceID kernels)
        c = node.value
        if isinstance(c, ast.Call) and isinstance(c.func, ast.Attribute) and c.func.attr == "mesh_query_ray":
            _Rewrite.count += 1
            q = f"_q{_Rewrite.count}"
            call = ast.Call(func=ast.Attribute(value=ast.Name("wp", ast.Load()), attr="_mesh_query_ray", ctx=ast.Load()), args=c.args[:4], keywords=[])
            hit = ast.Subscript(ast.Name(q, ast.Load()), ast.Constant(0), ast.Load())
            outs = ast.Tuple([ast.Name(a.id, ast.Store()) for a in c.args[4:10]], ast.Store())
            take = ast.If(test=hit, body=[ast.Assign([outs], ast.Subscript(ast.Name(q, ast.Load()), ast.Slice(ast.Constant(1), None, None), ast.Load()))], orelse=[])
            return [ast.Assign([ast.Name(q, ast.Store())], call), take]
        return node


def kernel(fn):
    src = textwrap.dedent(inspect.getsource(fn))
    tree = ast.parse(src)
    fd = tree.body[0]
    fd.decorator_list = []
    for a in fd.args.args:
        a.annotation = None
    tree = ast.fix_missing_locations(_Rewrite().visit(tree))
    ns = dict(fn.__globals__)
    ns["wp"] = wp
    exec(compile(tree, inspect.getsourcefile(fn), "exec"), ns)
    return ns[fd.name]


wp.kernel = kernel


def launch(kernel, dim, inputs, device=None, **kw):
    global _tid
    for idx in np.ndindex(*dim):
        _tid = idx
        kernel(*inputs)


wp.launch = launch


def install():
    sys.modules["warp"] = wp
    return wp
