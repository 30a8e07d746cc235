"""ctypes bindings of libarn_b200.so — the C-ABI in include/arn.h and include/arn_host.h.

This is plumbing for the Python harness (tests/, bench.py, __graft_entry__.py).  There is no
Python or CPU implementation behind these calls: if the shared library is missing the import
fails loudly, and every compute entry point needs a CUDA device.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ARN_LIB_PATH") or os.path.join(_HERE, "libarn_b200.so")   # ARN_LIB_PATH: A/B builds of the same library (tools/)

ARN_OK, ARN_E_INVALID, ARN_E_CUDA, ARN_E_OOM, ARN_E_IO, ARN_E_UNSUPPORTED, ARN_E_NCCL = 0, -1, -2, -3, -4, -5, -6
ARN_PRIM_SPHERE = 0x80000000
ARN_BVH_SAH, ARN_BVH_MIDDLECOUNT, ARN_BVH_MIDPOINT = 0, 1, 2
ARN_OPT_COUNT_TRAVERSAL, ARN_OPT_WAVE_CAPACITY, ARN_OPT_BVH_WIDTH, ARN_OPT_PIPELINES, ARN_OPT_TRACE_REFILL, ARN_OPT_SMEM_NODES, ARN_OPT_PDL = 1, 2, 3, 4, 5, 6, 7
ARN_SMEM_NODE_BYTES = 160 * 1024        # arn.h: budget of the pair records k_trace stages in shared memory (128 B per interior node)
ARN_LIGHT_POINT, ARN_LIGHT_SPOT, ARN_LIGHT_DISTANT = 0, 1, 2
ARN_LIGHT_ANALYTIC = 0x80000000
ARN_FILTER_LANCZOS, ARN_FILTER_BOX, ARN_FILTER_TRIANGLE, ARN_FILTER_GAUSSIAN, ARN_FILTER_MITCHELL = 0, 1, 2, 3, 4
ARN_MAT_MATTE, ARN_MAT_PLASTIC, ARN_MAT_GLASS, ARN_MAT_TRANSLUCENT = 0, 1, 2, 3
ARN_WRAP_REPEAT, ARN_WRAP_BLACK, ARN_WRAP_CLAMP = 0, 1, 2
ARN_SAMPLER_PARITY, ARN_SAMPLER_STRATIFIED = 0, 1

c_float_p = C.POINTER(C.c_float)
c_u32_p = C.POINTER(C.c_uint32)


class Node(C.Structure):
    _fields_ = [("bmin", C.c_float * 3), ("bmax", C.c_float * 3), ("offset", C.c_uint32), ("len_axis", C.c_uint32)]


class Material(C.Structure):
    _fields_ = [("type", C.c_uint32), ("kd", C.c_float * 3), ("ks", C.c_float * 3), ("sigma", C.c_float),
                ("roughness", C.c_float), ("alpha", C.c_float), ("eta", C.c_float), ("dissolve", C.c_float),
                ("kd_tex", C.c_uint32), ("ks_tex", C.c_uint32), ("aux_tex", C.c_uint32), ("bump_tex", C.c_uint32)]


class Texture(C.Structure):
    """arn_texture: an ImageTexture's mip pyramid + UVMapping (include/arn.h)."""
    _fields_ = [("channels", C.c_uint32), ("n_levels", C.c_uint32), ("trilinear", C.c_uint32), ("wrapping", C.c_uint32), ("max_aniso", C.c_float),
                ("scale_u", C.c_float), ("scale_v", C.c_float), ("shift_u", C.c_float), ("shift_v", C.c_float),
                ("level_w", C.c_uint32 * 16), ("level_h", C.c_uint32 * 16), ("level_offset", C.c_uint32 * 16)]


class Mesh(C.Structure):
    _fields_ = [("material", C.c_uint32), ("has_normals", C.c_uint32), ("has_uvs", C.c_uint32), ("reserved", C.c_uint32)]


class Sphere(C.Structure):
    _fields_ = [("radius", C.c_float), ("zmin", C.c_float), ("zmax", C.c_float), ("phimax", C.c_float),
                ("thetamin", C.c_float), ("thetamax", C.c_float), ("material", C.c_uint32),
                ("has_transform", C.c_uint32), ("emissive", C.c_uint32), ("emission", C.c_float * 3),
                ("local_parent", C.c_float * 16), ("parent_local", C.c_float * 16)]


class AnalyticLight(C.Structure):
    """arn_analytic_light: PointLight / SpotLight / DistantLight (include/arn.h)."""
    _fields_ = [("type", C.c_uint32), ("pos", C.c_float * 3), ("intensity", C.c_float * 3), ("cost", C.c_float), ("cosf", C.c_float),
                ("parent_local", C.c_float * 16), ("dir", C.c_float * 3), ("world_radius", C.c_float)]


class SceneDesc(C.Structure):
    _fields_ = [("n_vertices", C.c_uint32), ("positions", c_float_p), ("normals", c_float_p), ("uvs", c_float_p),
                ("n_triangles", C.c_uint32), ("indices", c_u32_p), ("tri_mesh", c_u32_p),
                ("n_meshes", C.c_uint32), ("meshes", C.POINTER(Mesh)),
                ("n_spheres", C.c_uint32), ("spheres", C.POINTER(Sphere)),
                ("n_materials", C.c_uint32), ("materials", C.POINTER(Material)),
                ("n_prims", C.c_uint32), ("prims", c_u32_p),
                ("n_nodes", C.c_uint32), ("nodes", C.POINTER(Node)), ("order", c_u32_p),
                ("n_lights", C.c_uint32), ("light_prims", c_u32_p), ("light_func", c_float_p), ("light_cdf", c_float_p),
                ("light_func_integral", C.c_float),
                ("n_analytic_lights", C.c_uint32), ("analytic_lights", C.POINTER(AnalyticLight)),
                ("n_textures", C.c_uint32), ("textures", C.POINTER(Texture)), ("n_texel_floats", C.c_uint64), ("texels", c_float_p)]


class Camera(C.Structure):
    _fields_ = [("raster_view", C.c_float * 16), ("view_parent", C.c_float * 16), ("has_lens", C.c_uint32),
                ("lens_radius", C.c_float), ("focal_distance", C.c_float), ("ortho", C.c_uint32)]


class Film(C.Structure):
    _fields_ = [("res_x", C.c_uint32), ("res_y", C.c_uint32), ("crop_min_x", C.c_int32), ("crop_min_y", C.c_int32),
                ("crop_max_x", C.c_int32), ("crop_max_y", C.c_int32), ("filter_radius_x", C.c_float), ("filter_radius_y", C.c_float),
                ("filter_kind", C.c_uint32), ("filter_a", C.c_float), ("filter_b", C.c_float)]


class Sampler(C.Structure):
    _fields_ = [("sampledx", C.c_uint32), ("sampledy", C.c_uint32), ("ndim", C.c_uint32), ("seed", C.c_uint32), ("mode", C.c_uint32)]


class PTParams(C.Structure):
    _fields_ = [("max_depth", C.c_uint32), ("min_depth", C.c_uint32), ("rr_threshold", C.c_float),
                ("tiles_x", C.c_uint32), ("tiles_y", C.c_uint32), ("rank", C.c_uint32), ("world_size", C.c_uint32),
                ("spp_begin", C.c_uint32), ("spp_end", C.c_uint32), ("partition_subdiv", C.c_uint32)]


class Ray(C.Structure):
    _fields_ = [("o", C.c_float * 3), ("d", C.c_float * 3), ("tmax", C.c_float)]


class Hit(C.Structure):
    _fields_ = [("prim_id", C.c_int32), ("t", C.c_float)]


class Stats(C.Structure):
    _fields_ = [("camera_rays", C.c_uint64), ("extend_rays", C.c_uint64), ("shadow_rays", C.c_uint64),
                ("mis_rays", C.c_uint64), ("invalid_samples", C.c_uint64), ("kernel_launches", C.c_uint64),
                ("gpu_ms", C.c_double), ("extend_ms", C.c_double), ("extend_bounce_ms", C.c_double),
                ("extend_bounce_rays", C.c_uint64), ("extend_nodes", C.c_uint64), ("extend_tris", C.c_uint64),
                ("extend_spheres", C.c_uint64)]


# every symbol include/arn.h and include/arn_host.h declare (checked by tests/test_abi.py)
ARN_H_SYMBOLS = [
    "arn_bvh_build", "arn_light_distribution", "arn_film_finalize", "arn_ctx_create", "arn_ctx_destroy",
    "arn_last_error", "arn_scene_upload", "arn_scene_destroy", "arn_intersect_closest", "arn_intersect_any",
    "arn_intersect_closest_dev", "arn_intersect_any_dev", "arn_intersect_closest_counted_dev",
    "arn_bvh_build_gpu", "arn_render_pt", "arn_render_pt_dev", "arn_render_pt_samples", "arn_ctx_set_option", "arn_ctx_synchronize", "arn_ctx_stream", "arn_version",
    "arn_film_reduce", "arn_film_merge", "arn_nccl_unique_id", "arn_nccl_comm_create", "arn_nccl_comm_destroy", "arn_selftest_math", "arn_selftest_bsdf", "arn_selftest_pair_records",
]
ARN_HOST_H_SYMBOLS = [
    "arn_hscene_create", "arn_hscene_destroy", "arn_hscene_last_error", "arn_hscene_add_material",
    "arn_hscene_add_mesh", "arn_hscene_add_sphere", "arn_hscene_add_light", "arn_hscene_add_texture", "arn_hscene_add_texture_file", "arn_spot_light_make", "arn_point_light_make",
    "arn_distant_light_make", "arn_hscene_load_obj", "arn_hscene_load_json",
    "arn_hscene_build", "arn_hscene_build_gpu", "arn_hscene_desc", "arn_camera_make", "arn_ortho_camera_make", "arn_save_png",
]

_lib = None


def load():
    """Load libarn_b200.so (built by __graft_entry__.build() / arendur_b200/csrc/Makefile)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no Python/CPU fallback for the path-tracing core)")
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    vp = C.c_void_p
    sig = {
        "arn_bvh_build": (C.c_int, [C.c_uint32, c_float_p, c_float_p, C.c_int, C.POINTER(Node), c_u32_p, c_u32_p]),
        "arn_light_distribution": (C.c_int, [C.c_uint32, c_float_p, c_float_p, c_float_p]),
        "arn_film_finalize": (C.c_int, [vp, C.c_size_t, vp, vp]),
        "arn_ctx_create": (C.c_int, [C.c_int, C.POINTER(vp)]),
        "arn_ctx_destroy": (None, [vp]),
        "arn_last_error": (C.c_char_p, [vp]),
        "arn_scene_upload": (C.c_int, [vp, C.POINTER(SceneDesc), C.POINTER(vp)]),
        "arn_scene_destroy": (None, [vp]),
        "arn_intersect_closest": (C.c_int, [vp, vp, C.c_size_t, vp]),
        "arn_intersect_any": (C.c_int, [vp, vp, C.c_size_t, vp]),
        "arn_intersect_closest_dev": (C.c_int, [vp, vp, C.c_size_t, vp, C.POINTER(Stats)]),
        "arn_intersect_any_dev": (C.c_int, [vp, vp, C.c_size_t, vp, C.POINTER(Stats)]),
        "arn_intersect_closest_counted_dev": (C.c_int, [vp, vp, C.c_size_t, vp, C.POINTER(C.c_uint64)]),
        "arn_render_pt": (C.c_int, [vp, C.POINTER(Camera), C.POINTER(Film), C.POINTER(Sampler), C.POINTER(PTParams), vp, C.POINTER(Stats)]),
        "arn_render_pt_dev": (C.c_int, [vp, C.POINTER(Camera), C.POINTER(Film), C.POINTER(Sampler), C.POINTER(PTParams), vp, C.POINTER(Stats)]),
        "arn_render_pt_samples": (C.c_int, [vp, C.POINTER(Camera), C.POINTER(Film), C.POINTER(Sampler), C.POINTER(PTParams), vp, vp, C.POINTER(Stats)]),
        "arn_ctx_set_option": (C.c_int, [vp, C.c_int, C.c_longlong]),
        "arn_ctx_synchronize": (C.c_int, [vp]),
        "arn_ctx_stream": (vp, [vp]),
        "arn_version": (C.c_char_p, []),
        "arn_film_reduce": (C.c_int, [vp, vp, vp, C.c_size_t, C.c_int]),
        "arn_film_merge": (C.c_int, [vp, vp, vp, C.c_size_t]),
        "arn_nccl_unique_id": (C.c_int, [vp]),
        "arn_nccl_comm_create": (C.c_int, [vp, vp, C.c_int, C.c_int, C.POINTER(vp)]),
        "arn_nccl_comm_destroy": (C.c_int, [vp]),
        "arn_selftest_math": (C.c_int, [vp, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint64)]),
        "arn_selftest_pair_records": (C.c_int, [vp, C.c_uint32, vp, C.POINTER(C.c_uint32)]),
        "arn_selftest_bsdf": (C.c_int, [vp, C.POINTER(Material), C.c_size_t, vp, vp, vp, vp, vp]),
        "arn_hscene_create": (C.c_int, [C.POINTER(vp)]),
        "arn_hscene_destroy": (None, [vp]),
        "arn_hscene_last_error": (C.c_char_p, [vp]),
        "arn_hscene_add_material": (C.c_int, [vp, C.POINTER(Material)]),
        "arn_hscene_add_mesh": (C.c_int, [vp, vp, C.c_uint32, vp, C.c_uint32, vp, vp, vp, C.c_uint32]),
        "arn_hscene_add_sphere": (C.c_int, [vp, C.c_float, C.c_float, C.c_float, C.c_float, C.c_uint32, vp, vp]),
        "arn_hscene_add_light": (C.c_int, [vp, C.POINTER(AnalyticLight)]),
        "arn_hscene_add_texture": (C.c_int, [vp, C.POINTER(Texture), vp, C.c_uint64]),
        "arn_hscene_add_texture_file": (C.c_int, [vp, C.c_char_p, C.POINTER(Texture), C.c_int, C.c_float, vp]),
        "arn_spot_light_make": (C.c_int, [vp, vp, vp, C.c_float, C.c_float, C.POINTER(AnalyticLight)]),
        "arn_point_light_make": (C.c_int, [vp, vp, C.POINTER(AnalyticLight)]),
        "arn_distant_light_make": (C.c_int, [vp, vp, C.c_float, C.POINTER(AnalyticLight)]),
        "arn_hscene_load_obj": (C.c_int, [vp, C.c_char_p, vp]),
        "arn_hscene_load_json": (C.c_int, [vp, C.c_char_p, C.c_char_p, C.POINTER(Camera), C.POINTER(Film), C.POINTER(Sampler), C.POINTER(PTParams), C.c_char_p, C.c_size_t]),
        "arn_hscene_build": (C.c_int, [vp, C.c_int]),
        "arn_hscene_build_gpu": (C.c_int, [vp, vp, C.POINTER(C.c_float)]),
        "arn_bvh_build_gpu": (C.c_int, [vp, C.c_uint32, vp, vp, vp, C.POINTER(C.c_uint32), C.POINTER(C.c_float)]),
        "arn_hscene_desc": (C.POINTER(SceneDesc), [vp]),
        "arn_camera_make": (C.c_int, [vp, vp, C.c_float, C.c_float, C.c_float, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, C.POINTER(Camera)]),
        "arn_ortho_camera_make": (C.c_int, [vp, vp, C.c_float, C.c_float, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, C.POINTER(Camera)]),
        "arn_save_png": (C.c_int, [C.c_char_p, vp, C.c_uint32, C.c_uint32]),
    }
    for name, (res, args) in sig.items():
        try:
            fn = getattr(lib, name)
        except AttributeError:
            if os.environ.get("ARN_LIB_PATH"):        # an older A/B build (tools/ab_trace.py) may lack newer entry points
                continue
            raise
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
