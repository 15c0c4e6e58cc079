"""ctypes binding of the C ABI declared in include/sightpy_b200.h.

This is the *only* way the package renders: if ``libsightpy_b200.so`` is missing or no CUDA
device can be bound, importing/using it raises — there is no CPU fallback.
"""
import ctypes as C
import os
from pathlib import Path

import numpy as np

from .flatten import (CAMERA_DT, COLLIDER_DT, LIGHT_DT, MATERIAL_DT, PRIMITIVE_DT, FlatScene)

__all__ = ["NativeScene", "NativeGroup", "load_library", "library_path", "Stats", "measure_peaks", "visible_devices",
           "configured_devices", "trim"]

SP_MAX_DEPTH_LEVELS = 64
_LIB = None
_BOUND_DEVICE = None


class Stats(C.Structure):
    _fields_ = [
        ("rays_total", C.c_uint64), ("shadow_rays", C.c_uint64),
        ("rays_per_depth", C.c_uint64 * SP_MAX_DEPTH_LEVELS),
        ("kernel_launches", C.c_uint64), ("chunks", C.c_uint64),
        ("device_ms", C.c_double), ("level_kernel_ms", C.c_double),
        ("level_kernel_launches", C.c_uint64), ("queue_bytes", C.c_uint64),
        ("level_ms", C.c_double * SP_MAX_DEPTH_LEVELS),
        ("peak_ray_records", C.c_uint64), ("peak_fan_records", C.c_uint64),
        ("warp_kernel_launches", C.c_uint64), ("chunk_retries", C.c_uint64),
    ]

    def as_dict(self):
        depth = [int(v) for v in self.rays_per_depth]
        while depth and depth[-1] == 0:
            depth.pop()
        return dict(rays_total=int(self.rays_total), shadow_rays=int(self.shadow_rays),
                    rays_per_depth=depth, kernel_launches=int(self.kernel_launches),
                    chunks=int(self.chunks), device_ms=float(self.device_ms),
                    level_kernel_ms=float(self.level_kernel_ms),
                    level_kernel_launches=int(self.level_kernel_launches),
                    queue_bytes=int(self.queue_bytes),
                    level_ms=[float(v) for v in self.level_ms][:max(len(depth), 1)],
                    peak_ray_records=int(self.peak_ray_records),
                    peak_fan_records=int(self.peak_fan_records),
                    warp_kernel_launches=int(self.warp_kernel_launches),
                    chunk_retries=int(self.chunk_retries))


def library_path():
    override = os.environ.get("SIGHTPY_B200_LIB")
    if override:
        return Path(override)
    return Path(__file__).resolve().parent.parent / "csrc" / "libsightpy_b200.so"


def load_library():
    """dlopen the CUDA library and declare its prototypes.  Raises if it was not built."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not path.exists():
        raise RuntimeError(
            f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C python-raytracer_b200/csrc`). sightpy-b200 has no CPU fallback.")
    lib = C.CDLL(str(path))
    vp, i32, u64, f64p = C.c_void_p, C.c_int, C.c_uint64, C.POINTER(C.c_double)
    proto = {
        "sp_abi_version": (i32, []),
        "sp_abi_sizes": (i32, [vp]),
        "sp_init": (i32, [i32]),
        "sp_init_devices": (i32, [i32, vp]),
        "sp_default_device": (i32, []),
        "sp_trim": (None, []),
        "sp_scene_create_on": (i32, [C.POINTER(vp), i32]),
        "sp_render_group": (i32, [vp, i32, i32, u64, i32, vp, vp, C.POINTER(Stats)]),
        "sp_device_count": (i32, []),
        "sp_last_error": (C.c_char_p, []),
        "sp_shutdown": (None, []),
        "sp_scene_create": (i32, [C.POINTER(vp)]),
        "sp_scene_destroy": (None, [vp]),
        "sp_scene_set_globals": (i32, [vp, vp, vp, vp, i32]),
        "sp_scene_set_camera": (i32, [vp, vp]),
        "sp_scene_add_texture": (i32, [vp, vp, i32, i32, i32, C.POINTER(i32)]),
        "sp_scene_add_texture_keyed": (i32, [vp, u64, vp, i32, i32, i32, C.POINTER(i32)]),
        "sp_scene_add_texture_blurred": (i32, [vp, u64, vp, i32, i32, i32, C.c_double, C.POINTER(i32)]),
        "sp_scene_read_texture": (i32, [vp, i32, vp]),
        "sp_scene_set_materials": (i32, [vp, vp, i32]),
        "sp_scene_set_primitives": (i32, [vp, vp, i32]),
        "sp_scene_set_colliders": (i32, [vp, vp, i32]),
        "sp_scene_set_lights": (i32, [vp, vp, i32]),
        "sp_scene_set_importance": (i32, [vp, vp, i32]),
        "sp_scene_set_shadow_colliders": (i32, [vp, vp, i32]),
        "sp_scene_commit": (i32, [vp]),
        "sp_scene_clear_textures": (i32, [vp]),
        "sp_scene_update_camera": (i32, [vp, vp]),
        "sp_render": (i32, [vp, i32, u64, vp, vp, C.POINTER(Stats)]),
        "sp_render_samples": (i32, [vp, i32, i32, u64, i32, C.POINTER(Stats)]),
        "sp_render_region": (i32, [vp, C.c_int64, C.c_int64, i32, i32, u64, i32, C.POINTER(Stats)]),
        "sp_render_tiles": (i32, [vp, vp, i32, i32, i32, i32, u64, i32, C.POINTER(Stats)]),
        "sp_accum_device_ptr": (vp, [vp]),
        "sp_accum_bytes": (u64, [vp]),
        "sp_resolve": (i32, [vp, i32, vp, vp]),
        "sp_scene_set_stream": (i32, [vp, vp, i32]),
        "sp_trace": (i32, [vp, vp, vp, i32, u64, vp, vp, vp, C.POINTER(Stats)]),
        "sp_camera_rays": (i32, [vp, i32, u64, vp, vp]),
        "sp_distances": (i32, [vp, u64, vp]),
        "sp_aovs": (i32, [vp, i32, u64, vp, vp, vp]),
        "sp_set_option": (i32, [vp, C.c_char_p, C.c_int64]),
        "sp_measure_peaks": (i32, [f64p, f64p]),
    }
    for name, (res, args) in proto.items():
        fn = getattr(lib, name)       # AttributeError here == header/library mismatch
        fn.restype, fn.argtypes = res, args
    sizes = (C.c_int32 * 6)()
    lib.sp_abi_sizes(sizes)
    expect = [CAMERA_DT.itemsize, MATERIAL_DT.itemsize, PRIMITIVE_DT.itemsize,
              COLLIDER_DT.itemsize, LIGHT_DT.itemsize, C.sizeof(Stats)]
    if list(sizes) != expect:
        raise RuntimeError(f"ABI struct size mismatch: library {list(sizes)} vs binding {expect}")
    _LIB = lib
    return lib


def _check(lib, rc, what):
    if rc != 0:
        msg = lib.sp_last_error()
        raise RuntimeError(f"{what} failed: {msg.decode() if msg else 'unknown error'}")


def default_device():
    for var in ("SIGHTPY_DEVICE", "LOCAL_RANK"):
        if os.environ.get(var, "") != "":
            return int(os.environ[var])
    return 0


_BOUND = set()          # devices initialised so far


def bind_device(device=None):
    """Initialise one CUDA device (the first one bound becomes the library's default device)."""
    global _BOUND_DEVICE
    lib = load_library()
    device = default_device() if device is None else int(device)
    if device not in _BOUND:
        if _BOUND_DEVICE is None:
            _check(lib, lib.sp_init(device), f"sp_init({device})")
            _BOUND_DEVICE = device
            import atexit
            atexit.register(lib.sp_shutdown)          # return pooled device memory, streams and events
        else:                                         # a further device: keep the default, add this one
            ids = (C.c_int * (len(_BOUND) + 1))(_BOUND_DEVICE, *sorted(_BOUND - {_BOUND_DEVICE}), device)
            _check(lib, lib.sp_init_devices(len(ids), ids), f"sp_init_devices(+{device})")
        _BOUND.add(device)
    return lib


def visible_devices():
    return list(range(load_library().sp_device_count()))


def configured_devices():
    """Devices a Scene renders on.  SIGHTPY_DEVICES = "all" | "0,1,2" | unset.  Unset: every visible GPU when the
    process is not one rank of a torchrun job (WORLD_SIZE <= 1), as the reference's render() uses every core
    (scene.py:80); under torchrun each rank drives the one GPU named by LOCAL_RANK."""
    spec = os.environ.get("SIGHTPY_DEVICES", "").strip().lower()
    if int(os.environ.get("WORLD_SIZE", "1")) > 1 or os.environ.get("SIGHTPY_DEVICE", "") != "":
        return [default_device()]
    if spec in ("", "all"):
        n = load_library().sp_device_count()
        return list(range(n)) if n > 0 else [0]
    return [int(x) for x in spec.split(",") if x.strip() != ""]


def trim():
    """Return idle pooled device memory to the driver (sp_trim)."""
    load_library().sp_trim()


def measure_peaks(device=None):
    lib = bind_device(device)
    a, b = C.c_double(), C.c_double()
    _check(lib, lib.sp_measure_peaks(C.byref(a), C.byref(b)), "sp_measure_peaks")
    return dict(fp32_tflops=a.value, copy_gbs=b.value)


def _ptr(arr):
    return arr.ctypes.data_as(C.c_void_p) if arr is not None else None


class NativeScene:
    """A committed scene living on one GPU."""

    def __init__(self, flat: FlatScene, device=None):
        self.lib = bind_device(device)
        # the CUDA device all of this scene's memory and kernels live on
        self.device_index = default_device() if device is None else int(device)
        self.flat = flat
        self.width = int(flat.camera["width"])
        self.height = int(flat.camera["height"])
        self.handle = C.c_void_p()
        _check(self.lib, self.lib.sp_scene_create_on(C.byref(self.handle), self.device_index), "sp_scene_create")
        try:
            self._upload(flat)
        except Exception:
            self.close()
            raise

    TABLES = ("materials", "primitives", "colliders", "lights", "importance", "shadow_colliders")

    def update(self, flat):
        """Re-describe this committed scene in place (animation frames, in-place edits of a Scene): only what changed
        is handed over again.  A camera move alone needs no commit at all; anything else is re-committed on the same
        handle, which keeps the wavefront queues, the frame buffers, keyed textures and — while the scene keeps its
        shape — the queue-occupancy estimates of earlier frames.  Returns what was done."""
        old, lib, h = self.flat, self.lib, self.handle
        same_tex = len(old.textures) == len(flat.textures) and all(
            getattr(a, "key", 0) == getattr(b, "key", 1) and a.decode == b.decode for a, b in zip(old.textures, flat.textures))
        same_tables = (all(np.array_equal(getattr(old, k), getattr(flat, k)) for k in self.TABLES)
                       and np.array_equal(old.media, flat.media) and np.array_equal(old.ambient, flat.ambient))
        same_size = (int(old.camera["width"]), int(old.camera["height"])) == (int(flat.camera["width"]), int(flat.camera["height"]))
        if same_tex and same_tables and same_size:
            if old.camera.tobytes() == flat.camera.tobytes():
                return "unchanged"
            cam = np.ascontiguousarray(flat.camera.reshape(1))
            _check(lib, lib.sp_scene_update_camera(h, _ptr(cam)), "sp_scene_update_camera")
            self.flat = flat
            return "camera"
        if not same_tex:
            _check(lib, lib.sp_scene_clear_textures(h), "sp_scene_clear_textures")
        self._upload(flat, textures=not same_tex)
        self.flat = flat
        self.width, self.height = int(flat.camera["width"]), int(flat.camera["height"])
        return "commit"

    def _upload(self, flat, textures=True):
        lib, h = self.lib, self.handle
        amb = np.ascontiguousarray(flat.ambient, dtype=np.float64)
        mre = np.ascontiguousarray(flat.media.real, dtype=np.float64)
        mim = np.ascontiguousarray(flat.media.imag, dtype=np.float64)
        _check(lib, lib.sp_scene_set_globals(h, _ptr(amb), _ptr(mre), _ptr(mim), len(flat.media)), "set_globals")
        cam = np.ascontiguousarray(flat.camera.reshape(1))
        _check(lib, lib.sp_scene_set_camera(h, _ptr(cam)), "set_camera")
        for t in (flat.textures if textures else ()):
            tid = C.c_int(-1)
            # a sky box that wants blurring hands over its raw texels: the device blurs them (sp_imaging.cu)
            blur = float(getattr(t, "cube_blur", 0.0))
            u8 = np.ascontiguousarray(t.source_u8 if blur else t.u8)
            _check(lib, lib.sp_scene_add_texture_blurred(h, int(getattr(t, "key", 0)), _ptr(u8), u8.shape[0], u8.shape[1],
                                                         t.decode, blur, C.byref(tid)), "add_texture")
        for name, fn in (("materials", lib.sp_scene_set_materials), ("primitives", lib.sp_scene_set_primitives),
                         ("colliders", lib.sp_scene_set_colliders), ("lights", lib.sp_scene_set_lights),
                         ("importance", lib.sp_scene_set_importance),
                         ("shadow_colliders", lib.sp_scene_set_shadow_colliders)):
            arr = np.ascontiguousarray(getattr(flat, name))
            _check(lib, fn(h, _ptr(arr), len(arr)), "set_" + name)
        _check(lib, lib.sp_scene_commit(h), "sp_scene_commit")

    # ------------------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "handle", None) is not None and self.handle.value:
            self.lib.sp_scene_destroy(self.handle)
            self.handle = C.c_void_p()

    __del__ = close

    def set_option(self, name, value):
        _check(self.lib, self.lib.sp_set_option(self.handle, name.encode(), int(value)), f"set_option({name})")

    def render(self, spp, seed=0, want_linear=True):
        """Full frame on this GPU -> (uint8 H x W x 3, float32 3 x H x W linear or None, stats)."""
        srgb = np.empty((self.height, self.width, 3), dtype=np.uint8)
        lin = np.empty((3, self.height, self.width), dtype=np.float32) if want_linear else None
        st = Stats()
        _check(self.lib, self.lib.sp_render(self.handle, int(spp), int(seed), _ptr(lin), _ptr(srgb), C.byref(st)),
               "sp_render")
        return srgb, lin, st.as_dict()

    def render_samples(self, sample_begin, sample_end, seed=0, clear=True):
        st = Stats()
        _check(self.lib, self.lib.sp_render_samples(self.handle, int(sample_begin), int(sample_end), int(seed),
                                                    int(bool(clear)), C.byref(st)), "sp_render_samples")
        return st.as_dict()

    def render_region(self, pix_begin, pix_end, sample_begin, sample_end, seed=0, clear=True):
        """Accumulate samples [sample_begin, sample_end) of the pixels [pix_begin, pix_end) (row-major index)."""
        st = Stats()
        _check(self.lib, self.lib.sp_render_region(self.handle, int(pix_begin), int(pix_end), int(sample_begin),
                                                   int(sample_end), int(seed), int(bool(clear)), C.byref(st)),
               "sp_render_region")
        return st.as_dict()

    def render_tiles(self, tile_ids, tile_size, sample_begin, sample_end, seed=0, clear=True):
        """Accumulate samples [sample_begin, sample_end) of the listed square tiles (row-major tile ids)."""
        ids = np.ascontiguousarray(tile_ids, dtype=np.int32)
        st = Stats()
        _check(self.lib, self.lib.sp_render_tiles(self.handle, _ptr(ids), len(ids), int(tile_size), int(sample_begin),
                                                  int(sample_end), int(seed), int(bool(clear)), C.byref(st)),
               "sp_render_tiles")
        return st.as_dict()

    def n_tiles(self, tile_size):
        return -(-self.width // tile_size) * -(-self.height // tile_size)

    def accum_pointer(self):
        return int(self.lib.sp_accum_device_ptr(self.handle)), int(self.lib.sp_accum_bytes(self.handle))

    def set_stream(self, cuda_stream):
        """Run this scene's kernels on a caller-owned stream (int handle); None = library stream."""
        _check(self.lib, self.lib.sp_scene_set_stream(self.handle, C.c_void_p(cuda_stream or 0),
                                                      int(cuda_stream is not None)), "sp_scene_set_stream")

    def use_current_stream(self):
        """Enqueue on torch's current CUDA stream *of the library's device* (so that a following NCCL collective is
        ordered after the kernels).  torch's current device must be that device: a collective issued on another
        device's stream would not be ordered at all."""
        import torch
        if torch.cuda.current_device() != self.device_index:
            raise RuntimeError(f"torch's current CUDA device is {torch.cuda.current_device()} but sightpy is bound to device "
                               f"{self.device_index}: call torch.cuda.set_device({self.device_index}) (LOCAL_RANK) first")
        self.set_stream(torch.cuda.current_stream(self.device_index).cuda_stream)

    def accum_tensor(self):
        """torch.float32 view (no copy) of the device accumulation buffer (float4 per pixel)."""
        from .parallel import accum_as_tensor
        return accum_as_tensor(self)

    def resolve_on_device(self, spp_total):
        """Average + tonemap into the library's device buffers without copying the frame out."""
        _check(self.lib, self.lib.sp_resolve(self.handle, int(spp_total), None, None), "sp_resolve")

    def resolve(self, spp_total, want_linear=True):
        srgb = np.empty((self.height, self.width, 3), dtype=np.uint8)
        lin = np.empty((3, self.height, self.width), dtype=np.float32) if want_linear else None
        _check(self.lib, self.lib.sp_resolve(self.handle, int(spp_total), _ptr(lin), _ptr(srgb)), "sp_resolve")
        return srgb, lin

    def trace(self, origins, dirs, seed=0, want_rgb=True):
        o = np.ascontiguousarray(origins, dtype=np.float32)
        d = np.ascontiguousarray(dirs, dtype=np.float32)
        if o.shape != d.shape or o.ndim != 2 or o.shape[1] != 3:
            raise ValueError("origins/dirs must both be (n, 3)")
        n = o.shape[0]
        rgb = np.empty((n, 3), dtype=np.float32) if want_rgb else None
        hit = np.empty(n, dtype=np.int32)
        t = np.empty(n, dtype=np.float32)
        st = Stats()
        _check(self.lib, self.lib.sp_trace(self.handle, _ptr(o), _ptr(d), n, int(seed), _ptr(rgb), _ptr(hit),
                                           _ptr(t), C.byref(st)), "sp_trace")
        return dict(rgb=rgb, hit_id=hit, t=t, stats=st.as_dict())

    def camera_rays(self, sample=0, seed=0):
        n = self.width * self.height
        o = np.empty((n, 3), dtype=np.float32)
        d = np.empty((n, 3), dtype=np.float32)
        _check(self.lib, self.lib.sp_camera_rays(self.handle, int(sample), int(seed), _ptr(o), _ptr(d)),
               "sp_camera_rays")
        return o, d

    def read_texture(self, tex_id):
        """Texels of one texture as the device holds them (after any sky-box blur) -> H x W x 3 uint8."""
        t = self.flat.textures[tex_id]
        src = t.source_u8 if getattr(t, "cube_blur", 0.0) else t.u8
        out = np.empty(src.shape, dtype=np.uint8)
        _check(self.lib, self.lib.sp_scene_read_texture(self.handle, int(tex_id), _ptr(out)), "sp_scene_read_texture")
        return out

    def aovs(self, sample=0, seed=0):
        """Per-pixel nearest collider index, hit distance and ray-facing collider normal of one primary ray per pixel."""
        n = self.width * self.height
        hit = np.empty(n, dtype=np.int32)
        t = np.empty(n, dtype=np.float32)
        normal = np.empty((n, 3), dtype=np.float32)
        _check(self.lib, self.lib.sp_aovs(self.handle, int(sample), int(seed), _ptr(hit), _ptr(t), _ptr(normal)), "sp_aovs")
        return dict(hit_id=hit.reshape(self.height, self.width), t=t.reshape(self.height, self.width),
                    normal=normal.reshape(self.height, self.width, 3))

    def distances(self, seed=0):
        t = np.empty(self.width * self.height, dtype=np.float32)
        _check(self.lib, self.lib.sp_distances(self.handle, int(seed), _ptr(t)), "sp_distances")
        return t


class NativeGroup:
    """Replicas of one scene on several GPUs of the node, rendered together by sp_render_group: one host thread per
    device inside the library, frames gathered over NVLink peer access — no launcher, no torch (the reference's
    render() fans out over a process pool the same way, scene.py:98-116).  Everything that is not a frame render
    (trace, camera_rays, aovs, distances) goes to the first member."""
    GROUP_MIN_PRIMARIES = 4 << 20        # smaller frames are rendered by the first member alone

    def __init__(self, flat, devices):
        self.members = [NativeScene(flat, device=d) for d in devices]
        self.lib = self.members[0].lib

    # what Scene / parallel.render_frame use
    flat = property(lambda self: self.members[0].flat)
    width = property(lambda self: self.members[0].width)
    height = property(lambda self: self.members[0].height)
    device_index = property(lambda self: self.members[0].device_index)

    def update(self, flat):
        return [m.update(flat) for m in self.members][0]

    def close(self):
        for m in self.members:
            m.close()

    def render(self, spp, seed=0, want_linear=True, shard="auto"):
        first = self.members[0]
        if len(self.members) == 1 or int(spp) * self.width * self.height < self.GROUP_MIN_PRIMARIES:
            return first.render(spp, seed, want_linear)
        srgb = np.empty((self.height, self.width, 3), dtype=np.uint8)
        lin = np.empty((3, self.height, self.width), dtype=np.float32) if want_linear else None
        st = Stats()
        handles = (C.c_void_p * len(self.members))(*[m.handle for m in self.members])
        _check(self.lib, self.lib.sp_render_group(handles, len(self.members), int(spp), int(seed), 1 if shard == "tiles" else 0,
                                                  _ptr(lin), _ptr(srgb), C.byref(st)), "sp_render_group")
        out = st.as_dict()
        out["devices"] = [m.device_index for m in self.members]
        return srgb, lin, out

    def render_on_device(self, spp, seed=0, shard="samples"):
        """The same without copying the frame out (bench.py): the resolved frame stays in the first member's buffers."""
        st = Stats()
        handles = (C.c_void_p * len(self.members))(*[m.handle for m in self.members])
        _check(self.lib, self.lib.sp_render_group(handles, len(self.members), int(spp), int(seed), 1 if shard == "tiles" else 0,
                                                  None, None, C.byref(st)), "sp_render_group")
        return st.as_dict()

    def __getattr__(self, name):         # trace, camera_rays, aovs, distances, set_option, ...
        return getattr(self.members[0], name)
