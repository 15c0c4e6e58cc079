"""Scene geometry: user-facing primitives and the colliders they are made of.

Reference: sightpy/geometry/{primitive,collider,sphere,plane,cuboid,triangle,triangle_mesh}.py.
Constructor signatures and attribute names follow the reference so that scripts and the scene
flattener (flatten.py) see the same objects.  These classes only *describe* geometry: the
ray/collider intersection, normals and uv mapping run on the GPU (csrc/geometry.cuh), and the
float64 restatement used for testing lives in oracle/.
"""
import numpy as np

from .constants import *  # noqa: F401,F403  (re-exported like the reference does)
from .vec import vec3

__all__ = [
    "Primitive", "Collider", "Sphere", "Sphere_Collider", "Plane", "Plane_Collider",
    "Cuboid", "Cuboid_Collider", "Triangle", "Triangle_Collider", "TriangleMesh",
    "rotation_matrix",
]


def rotation_matrix(theta_deg, axis):
    """Rodrigues matrix as the reference builds it (primitive.py:15-44), including its
    ``sin = sqrt(1-cos^2)*sign(theta)`` shortcut (wrong beyond +-180 deg; kept for parity)."""
    u = axis.normalize()
    t = theta_deg / 180 * np.pi
    c = np.cos(t)
    s = np.sqrt(1 - c ** 2) * np.sign(t)
    k = 1 - c
    return np.array([
        [c + u.x * u.x * k,        u.x * u.y * k - u.z * s,  u.x * u.z * k + u.y * s],
        [u.y * u.x * k + u.z * s,  c + u.y ** 2 * k,         u.y * u.z * k - u.x * s],
        [u.z * u.x * k - u.y * s,  u.z * u.y * k + u.x * s,  c + u.z * u.z * k],
    ])


def _columns(a, b, c):
    """3x3 matrix whose columns are the vectors a, b, c."""
    return np.array([[a.x, b.x, c.x], [a.y, b.y, c.y], [a.z, b.z, c.z]])


class Primitive:
    """A renderable object: one material + a list of colliders (primitive.py:6-14)."""

    def __init__(self, center, material, max_ray_depth=1, shadow=True, mc=False):
        self.center = center
        self.material = material
        material.assigned_primitive = self
        self.shadow = shadow
        self.max_ray_depth = max_ray_depth
        self.mc = mc
        self.collider_list = []

    def rotate(self, θ, u):
        M = rotation_matrix(θ, u)
        for collider in self.collider_list:
            collider.rotate(M, self.center)

    # uv convention of the primitive; cuboid-like primitives rescale the cross layout
    uv_cross_layout = False

    def get_uv(self, hit):  # pragma: no cover - evaluated on the GPU
        raise NotImplementedError("uv mapping is evaluated by the CUDA backend")


class Collider:
    """Intersectable shape owned by a primitive (collider.py:7-18)."""

    def __init__(self, assigned_primitive, center):
        self.assigned_primitive = assigned_primitive
        self.center = center

    def rotate(self, M, center):
        pass


# ------------------------------------------------------------------------------------------------
class Sphere_Collider(Collider):
    def __init__(self, radius, **kwargs):
        super().__init__(**kwargs)
        self.radius = radius


class Sphere(Primitive):
    def __init__(self, center, material, radius, max_ray_depth=5, shadow=True, mc=False):
        super().__init__(center, material, max_ray_depth, shadow=shadow, mc=mc)
        self.collider_list.append(Sphere_Collider(assigned_primitive=self, center=center, radius=radius))
        self.bounded_sphere_radius = radius


# ------------------------------------------------------------------------------------------------
class Plane_Collider(Collider):
    """Bounded rectangle: |u_axis.(M-C)| <= w and |v_axis.(M-C)| <= h (plane.py:39-55)."""

    def __init__(self, u_axis, v_axis, w, h, uv_shift=(0.0, 0.0), **kwargs):
        super().__init__(**kwargs)
        self.u_axis, self.v_axis = u_axis, v_axis
        self.normal = u_axis.cross(v_axis).normalize()
        self.w, self.h = w, h
        self.uv_shift = uv_shift
        # tangent frame used by normal maps.  NOT refreshed by rotate() (plane.py:92-96).
        self.inverse_basis_matrix = _columns(self.u_axis, self.v_axis, self.normal)
        self.basis_matrix = self.inverse_basis_matrix.T

    def rotate(self, M, center):
        self.u_axis = self.u_axis.matmul(M)
        self.v_axis = self.v_axis.matmul(M)
        self.normal = self.normal.matmul(M)
        self.center = center + (self.center - center).matmul(M)


class Plane(Primitive):
    def __init__(self, center, material, width, height, u_axis, v_axis, max_ray_depth=5, shadow=True):
        super().__init__(center, material, max_ray_depth, shadow=shadow)
        self.collider_list.append(Plane_Collider(
            assigned_primitive=self, center=center, u_axis=u_axis, v_axis=v_axis,
            w=width / 2, h=height / 2))
        self.width, self.height = width, height
        self.bounded_sphere_radius = np.sqrt((width / 2) ** 2 + (height / 2) ** 2)


# ------------------------------------------------------------------------------------------------
class Cuboid_Collider(Collider):
    """Oriented box, slab-tested in its own basis (cuboid.py:60-103)."""

    def __init__(self, width, height, length, **kwargs):
        super().__init__(**kwargs)
        half = vec3(width / 2, height / 2, length / 2)
        self.lb = self.center - half
        self.rt = self.center + half
        self.lb_local_basis, self.rt_local_basis = self.lb, self.rt
        self.width, self.height, self.length = width, height, length
        self.ax_w, self.ax_h, self.ax_l = vec3(1.0, 0.0, 0.0), vec3(0.0, 1.0, 0.0), vec3(0.0, 0.0, 1.0)
        self._refresh_basis()

    def _refresh_basis(self):
        self.inverse_basis_matrix = _columns(self.ax_w, self.ax_h, self.ax_l)
        self.basis_matrix = self.inverse_basis_matrix.T

    def rotate(self, M, center):
        self.ax_w, self.ax_h, self.ax_l = (a.matmul(M) for a in (self.ax_w, self.ax_h, self.ax_l))
        self._refresh_basis()
        self.lb = center + (self.lb - center).matmul(M)
        self.rt = center + (self.rt - center).matmul(M)
        self.lb_local_basis = self.lb.matmul(self.basis_matrix)
        self.rt_local_basis = self.rt.matmul(self.basis_matrix)


class Cuboid(Primitive):
    uv_cross_layout = True   # (u/4, v/3) rescale of the cube-map cross (cuboid.py:29-32)

    def __init__(self, center, material, width, height, length, max_ray_depth=5, shadow=True):
        super().__init__(center, material, max_ray_depth, shadow=shadow)
        self.width, self.height, self.length = width, height, length
        self.bounded_sphere_radius = np.sqrt((width / 2) ** 2 + (height / 2) ** 2 + (length / 2) ** 2)
        self.collider_list.append(Cuboid_Collider(
            assigned_primitive=self, center=center, width=width, height=height, length=length))


# ------------------------------------------------------------------------------------------------
class Triangle_Collider(Collider):
    """Triangle with precomputed inward edge normals (triangle.py:20-35).  ``assigned_surface``
    is the reference's (inconsistent) keyword; ``assigned_primitive`` is accepted too."""

    def __init__(self, assigned_surface=None, p1=None, p2=None, p3=None, assigned_primitive=None):
        owner = assigned_surface if assigned_surface is not None else assigned_primitive
        self.p1, self.p2, self.p3 = p1, p2, p3
        self.normal = (p2 - p1).cross(p3 - p1).normalize()
        self.centroid = (p1 + p2 + p3) / 3
        super().__init__(assigned_primitive=owner, center=self.centroid)
        self.n31 = (p3 - p1).cross(self.normal)
        self.n12 = (p1 - p2).cross(self.normal)
        self.n23 = (p2 - p3).cross(self.normal)

    def rotate(self, M, center):
        self.p1, self.p2, self.p3 = (center + (p - center).matmul(M) for p in (self.p1, self.p2, self.p3))
        self.n31, self.n12, self.n23 = (n.matmul(M) for n in (self.n31, self.n12, self.n23))
        self.normal = self.normal.matmul(M)
        self.centroid = center + (self.centroid - center).matmul(M)
        self.center = self.centroid


def _bounding_radius(center, points):
    return float(max(np.sqrt((p - center).dot(p - center)) for p in points))


class Triangle(Primitive):
    """Single triangle.  The reference constructor cannot run (triangle.py:11-13 passes a keyword
    the collider does not take); this one keeps its signature and works.  Solid colours only."""

    def __init__(self, center, material, p1, p2, p3, max_ray_depth, shadow=True):
        super().__init__(center, material, max_ray_depth, shadow=shadow)
        self.collider_list.append(Triangle_Collider(assigned_primitive=self, p1=p1, p2=p2, p3=p3))
        self.bounded_sphere_radius = _bounding_radius(center, (p1, p2, p3))


class TriangleMesh(Primitive):
    """Wavefront-OBJ triangle soup, brute-force intersected (triangle_mesh.py:12-43, which
    raises NameError upstream; same signature, working)."""

    def __init__(self, file_name, center, material, max_ray_depth, shadow=True):
        super().__init__(center, material, max_ray_depth, shadow=shadow)
        vertices, faces = [], []
        with open(file_name, "r") as fh:
            for line in fh:
                tok = line.split()
                if not tok:
                    continue
                if tok[0] == "v":
                    vertices.append(vec3(float(tok[1]), float(tok[2]), float(tok[3])))
                elif tok[0] == "f":
                    faces.append([int(t.split("/")[0]) - 1 for t in tok[1:4]])
        for a, b, c in faces:
            self.collider_list.append(Triangle_Collider(
                assigned_primitive=self, p1=vertices[a] + center, p2=vertices[b] + center,
                p3=vertices[c] + center))
        pts = [v + center for v in vertices] or [center]
        self.bounded_sphere_radius = _bounding_radius(center, pts)
