"""ctypes loader for libtuch_b200.so (the C ABI in include/tuch_b200.h).

There is no CPU or PyTorch fallback: if the shared library is missing or a call fails, a
TuchError is raised."""
import ctypes as C
import os

from . import build as _build

_lib = None


class TuchError(RuntimeError):
    pass


class ContactFitArgs(C.Structure):
    """tuch_contact_fit_args of include/tuch_b200.h (field for field)."""
    _fields_ = ([(n, C.c_void_p) for n in (
        'body_pose', 'global_orient', 'exp_avg_pose', 'exp_avg_sq_pose', 'exp_avg_orient', 'exp_avg_sq_orient',
        'step_pose', 'step_orient', 'betas', 'camera_t', 'camera_center', 'joints_2d', 'joints_conf', 'body_active',
        'pair_active', 'smpl_workspace', 'vertices', 'joints', 'loss', 'per_body', 'exterior', 'argmin',
        'grad_body_pose', 'grad_global_orient')]
        + [(n, C.c_float) for n in ('euclthres', 'focal_length', 'sigma', 'pose_prior_weight', 'contact_loss_weight')]
        + [('use_segments', C.c_int)]
        + [(n, C.c_double) for n in ('lr', 'beta1', 'beta2', 'eps')])


def _declare(lib):
    vp, i32, f32 = C.c_void_p, C.c_int, C.c_float
    lib.tuch_last_error.restype = C.c_char_p
    lib.tuch_last_error.argtypes = []
    lib.tuch_abi_version.restype = i32
    lib.tuch_launch_count.restype = C.c_longlong
    lib.tuch_device_info.argtypes = [C.POINTER(i32)] * 3
    lib.tuch_release_scratch.argtypes = []
    lib.tuch_scratch_generation.argtypes = []
    lib.tuch_scratch_generation.restype = C.c_longlong
    lib.tuch_kernel_timing_enable.argtypes = [i32]
    lib.tuch_kernel_timing_reset.argtypes = []
    lib.tuch_kernel_timing_read.argtypes = [C.c_char_p, C.POINTER(C.c_double), C.POINTER(C.c_longlong)]
    lib.tuch_pairwise_dist.argtypes = [vp, vp, i32, i32, i32, i32, vp, vp]
    lib.tuch_pairwise_dist_backward.argtypes = [vp, vp, vp, vp, i32, i32, i32, i32, vp, vp, vp]
    lib.tuch_solid_angles.argtypes = [vp, vp, i32, i32, i32, vp, vp]
    lib.tuch_winding_numbers.argtypes = [vp, vp, i32, i32, i32, vp, vp]
    lib.tuch_topology_create.argtypes = [i32, i32, vp, C.POINTER(vp)]
    lib.tuch_topology_destroy.argtypes = [vp]
    lib.tuch_topology_destroy.restype = None
    lib.tuch_topology_num_verts.argtypes = [vp]
    lib.tuch_topology_strip_stats.argtypes = [vp, C.POINTER(i32), C.POINTER(i32)]
    lib.tuch_topology_set_template.argtypes = [vp, vp]
    lib.tuch_topology_set_winding_mode.argtypes = [vp, i32]
    lib.tuch_topology_cluster_stats.argtypes = [vp] + [C.POINTER(i32)] * 5
    lib.tuch_topology_query_stats.argtypes = [vp, C.POINTER(i32), C.POINTER(i32), vp]
    lib.tuch_topology_pack_nodes.argtypes = [vp, vp, i32, i32, vp, vp]
    lib.tuch_cluster_tree_host.argtypes = [vp, i32, i32, vp, vp, i32, vp, i32, vp, i32, vp, i32] + [C.POINTER(i32)] * 4
    lib.tuch_strip_stream_host.argtypes = [vp, i32, vp, vp, i32, C.POINTER(i32), C.POINTER(i32)]
    lib.tuch_topology_num_faces.argtypes = [vp]
    lib.tuch_topology_total_segment_verts.argtypes = [vp]
    lib.tuch_topology_set_geodist.argtypes = [vp, vp, f32, vp]
    lib.tuch_topology_set_geomask.argtypes = [vp, vp, vp]
    lib.tuch_topology_set_regions.argtypes = [vp, i32, vp, vp, i32, vp, vp]
    lib.tuch_topology_set_segments.argtypes = [vp, i32, vp, vp, vp, vp, vp, vp, vp]
    lib.tuch_contact_query.argtypes = [vp, vp, i32, i32, vp, vp, vp, vp, vp]
    lib.tuch_contact_query_within.argtypes = [vp, vp, i32, i32, f32, vp, vp, vp, vp, vp]
    lib.tuch_segment_exterior.argtypes = [vp, vp, i32, vp, vp, vp]
    lib.tuch_region_min.argtypes = [vp, vp, i32, i32, vp, vp, vp, vp, vp]
    lib.tuch_smpl_create.argtypes = [i32, i32, vp, vp, vp, vp, vp, vp, i32, vp, i32, vp, i32, vp, C.POINTER(vp)]
    lib.tuch_smpl_destroy.argtypes = [vp]
    lib.tuch_smpl_destroy.restype = None
    lib.tuch_smpl_num_verts.argtypes = [vp]
    lib.tuch_smpl_num_joints.argtypes = [vp]
    lib.tuch_smpl_workspace_floats.argtypes = [vp, i32]
    lib.tuch_smpl_workspace_floats.restype = C.c_size_t
    lib.tuch_smpl_forward.argtypes = [vp, vp, vp, i32, i32, vp, vp, vp, vp]
    lib.tuch_smpl_backward.argtypes = [vp, vp, i32, i32, vp, vp, vp, vp, vp, vp]
    lib.tuch_reprojection_loss.argtypes = [vp, vp, vp, vp, vp, i32, i32, f32, f32, vp, f32, vp, vp, vp, vp, vp, vp]
    lib.tuch_prior_create.argtypes = [i32, i32, vp, vp, vp, C.POINTER(vp)]
    lib.tuch_prior_destroy.argtypes = [vp]
    lib.tuch_prior_destroy.restype = None
    lib.tuch_pose_terms.argtypes = [vp, vp, vp, i32, i32, i32, f32, f32, f32, vp, vp, vp, vp, vp, vp]
    lib.tuch_contact_loss.argtypes = [vp, vp, vp, vp, vp, i32, i32, f32, i32, i32, f32, vp, vp, vp, vp, vp]
    lib.tuch_region_sum.argtypes = [vp, i32, i32, i32, vp, vp, vp, vp, f32, vp, vp, vp, vp]
    lib.tuch_contact_fit_step.argtypes = [vp, vp, vp, i32, C.POINTER(ContactFitArgs), vp]
    lib.tuch_adam_step.argtypes = [vp, vp, vp, vp, C.c_longlong, vp, C.c_double, C.c_double, C.c_double, C.c_double, vp]
    lib.tuch_topology_set_hd.argtypes = [vp, i32, vp, vp, vp, vp]
    lib.tuch_topology_num_hd.argtypes = [vp]
    lib.tuch_regressor_contact_loss.argtypes = [vp, vp, i32, vp, f32, i32, f32, vp, vp, vp, vp, vp, vp, vp, vp]
    lib.tuch_kernel_timing_names.argtypes = [C.c_char_p, i32]
    lib.tuch_estimate_translation.argtypes = [vp, vp, vp, i32, i32, i32, f32, f32, vp, vp]
    lib.tuch_rotmat_to_angle_axis.argtypes = [vp, i32, i32, vp, vp]
    lib.tuch_fits_pose_transform.argtypes = [vp, vp, vp, vp, i32, i32, i32, vp, vp]
    lib.tuch_winding_numbers_host.argtypes = [vp, vp, i32, i32, i32, vp]
    lib.tuch_contact_query_host.argtypes = [vp, vp, i32, i32, vp, vp, vp, vp]


def lib():
    global _lib
    if _lib is None:
        path = _build.LIB
        if not os.path.exists(path):
            raise TuchError('%s is missing: run `python -m tuch_b200.build` (or __graft_entry__.build()); '
                            'tuch_b200 has no CPU fallback' % path)
        _lib = C.CDLL(path)
        _declare(_lib)
    return _lib


def check(rc, what=''):
    if rc != 0:
        msg = lib().tuch_last_error().decode('utf-8', 'replace')
        raise TuchError('%s failed (%d): %s' % (what or 'tuch_b200 call', rc, msg))
