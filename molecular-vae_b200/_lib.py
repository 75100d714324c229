"""ctypes binding of libmvae_b200.so (the C ABI in include/mvae_b200.h).

There is no CPU or PyTorch fallback: if the CUDA library is missing the import fails loudly, and every
entry point raises when the library reports an error."""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libmvae_b200.so")

PREC_FP32 = 0
PREC_BF16 = 1


class CfgBDesc(ctypes.Structure):
    _fields_ = [
        ("batch", ctypes.c_int32), ("seq_len", ctypes.c_int32), ("charset", ctypes.c_int32),
        ("latent", ctypes.c_int32), ("hidden", ctypes.c_int32), ("layers", ctypes.c_int32),
        ("fc0", ctypes.c_int32), ("precision", ctypes.c_int32), ("train", ctypes.c_int32),
        ("max_len", ctypes.c_float), ("eps_scale", ctypes.c_float),
    ]


class CfgADesc(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in ("batch", "seq_len", "charset", "embed", "enc_hidden", "enc_layers", "latent",
                                              "dec_hidden", "dec_layers", "precision")] + [("max_len", ctypes.c_float),
                                                                                         ("eps_scale", ctypes.c_float)]


class MosesDesc(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in ("batch", "max_len", "vocab", "d_z", "q_hidden", "d_hidden", "d_layers",
                                              "mlp_hidden", "pad_id", "precision")] + [("kl_weight", ctypes.c_float),
                                                                                   ("recon_weight", ctypes.c_float),
                                                                                   ("q_bidir", ctypes.c_int32),
                                                                                   ("q_linear_heads", ctypes.c_int32),
                                                                                   ("d_dropout", ctypes.c_float),
                                                                                   ("dropout_seed", ctypes.c_uint32)]


class BindingDesc(ctypes.Structure):
    _fields_ = [("batch", ctypes.c_int32), ("z_size", ctypes.c_int32), ("train", ctypes.c_int32),
                ("bn_eps", ctypes.c_float), ("bn_momentum", ctypes.c_float)]


class MvaeError(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). molecular-vae_b200 has no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    vp, ll, i32 = ctypes.c_void_p, ctypes.c_longlong, ctypes.c_int
    pp = ctypes.POINTER(ctypes.c_void_p)
    dp = ctypes.POINTER(CfgBDesc)
    ap = ctypes.POINTER(CfgADesc)
    sigs = {
        "mvae_strerror": (ctypes.c_char_p, [i32]),
        "mvae_last_cuda_error": (ctypes.c_char_p, []),
        "mvae_launch_count": (ll, []),
        "mvae_reset_launch_count": (None, []),
        "mvae_profile_enable": (i32, [i32]),
        "mvae_profile_read": (i32, [i32, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(i32)]),
        "mvae_cfgb_workspace_bytes": (ctypes.c_size_t, [dp]),
        "mvae_cfgb_elbo_step": (i32, [dp, pp, pp, vp, vp, vp, vp, vp, vp, ctypes.c_size_t, vp]),
        "mvae_cfgb_elbo_step_graph_create": (i32, [dp, pp, pp, vp, vp, vp, vp, vp, vp, ctypes.c_size_t,
                                                   ctypes.POINTER(vp)]),
        "mvae_cfgb_elbo_step_phase": (i32, [dp, pp, pp, vp, vp, vp, vp, vp, vp, ctypes.c_size_t, i32, vp]),
        "mvae_cfgb_elbo_step_phase_graph_create": (i32, [dp, pp, pp, vp, vp, vp, vp, vp, vp, ctypes.c_size_t, i32,
                                                         ctypes.POINTER(vp)]),
        "mvae_graph_launch": (i32, [vp, vp]),
        "mvae_graph_num_kernel_nodes": (ll, [vp]),
        "mvae_graph_destroy": (None, [vp]),
        "mvae_cfgb_forward": (i32, [dp, pp, vp, vp, vp, vp, vp, vp, ctypes.c_size_t, vp]),
        "mvae_cfgb_backward": (i32, [dp, pp, pp, vp, vp, vp, vp, vp, vp, ctypes.c_size_t, vp]),
        "mvae_cfgb_decode_greedy": (i32, [dp, pp, vp, vp, vp, vp, ctypes.c_size_t, vp]),
        "mvae_onehot_to_ids": (i32, [vp, ll, i32, vp, vp, vp]),
        "mvae_cfgb_read_error": (i32, [dp, vp, ctypes.c_size_t, ctypes.POINTER(i32), vp]),
        "mvae_cfga_workspace_bytes": (ctypes.c_size_t, [ap]),
        "mvae_cfga_elbo_step": (i32, [ap, pp, pp, vp, vp, vp, vp, vp, vp, ctypes.c_size_t, vp]),
        "mvae_cfga_elbo_step_graph_create": (i32, [ap, pp, pp, vp, vp, vp, vp, vp, vp, ctypes.c_size_t,
                                                   ctypes.POINTER(vp)]),
        "mvae_cfga_forward": (i32, [ap, pp, vp, vp, vp, vp, vp, vp, ctypes.c_size_t, vp]),
        "mvae_cfga_backward": (i32, [ap, pp, pp, vp, vp, vp, vp, vp, vp, ctypes.c_size_t, vp]),
        "mvae_cfga_decode": (i32, [ap, pp, vp, vp, vp, vp, ctypes.c_size_t, vp]),
        "mvae_cfga_read_error": (i32, [ap, vp, ctypes.c_size_t, ctypes.POINTER(i32), vp]),
        "mvae_moses_workspace_bytes": (ctypes.c_size_t, [ctypes.POINTER(MosesDesc)]),
        "mvae_moses_step": (i32, [ctypes.POINTER(MosesDesc), pp, pp, vp, vp, vp, vp, vp, vp, vp, vp, vp, ctypes.c_size_t, vp]),
        "mvae_moses_step_ex": (i32, [ctypes.POINTER(MosesDesc), pp, pp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, ctypes.c_size_t, i32, vp]),
        "mvae_moses_joint_extra_bytes": (ctypes.c_size_t, [ctypes.POINTER(MosesDesc)]),
        "mvae_moses_joint_step": (i32, [ctypes.POINTER(MosesDesc), pp, pp, vp, vp, vp, vp, ctypes.POINTER(BindingDesc), pp, pp, pp, vp,
                                        ctypes.c_float, vp, vp, vp, vp, ctypes.c_size_t, vp, ctypes.c_size_t, vp, ctypes.c_size_t, i32, vp]),
        "mvae_moses_sample": (i32, [ctypes.POINTER(MosesDesc), pp, vp, i32, i32, i32, ctypes.c_float, ctypes.c_ulonglong, vp, vp, vp,
                                    vp, ctypes.c_size_t, vp]),
        "mvae_moses_sample_graph_create": (i32, [ctypes.POINTER(MosesDesc), pp, vp, i32, i32, i32, ctypes.c_float, vp, vp, vp,
                                                 vp, ctypes.c_size_t, ctypes.POINTER(vp)]),
        "mvae_text_to_ids": (i32, [vp, vp, i32, i32, vp, i32, vp, vp, vp]),
        "mvae_moses_read_error": (i32, [ctypes.POINTER(MosesDesc), vp, ctypes.c_size_t, ctypes.POINTER(i32), vp]),
        "mvae_binding_workspace_bytes": (ctypes.c_size_t, [ctypes.POINTER(BindingDesc)]),
        "mvae_binding_forward": (i32, [ctypes.POINTER(BindingDesc), pp, pp, vp, vp, vp, ctypes.c_size_t, vp]),
        "mvae_binding_backward": (i32, [ctypes.POINTER(BindingDesc), pp, pp, vp, vp, vp, vp, ctypes.c_size_t, vp]),
        "mvae_ids_to_text": (i32, [vp, vp, i32, i32, vp, i32, vp, i32, i32, i32, vp, ll, vp, vp, vp]),
        "mvae_clip_grad_norm": (i32, [vp, ll, ctypes.c_float, vp, vp, i32, vp]),
        "mvae_adam_step": (i32, [vp, vp, vp, vp, ll] + [ctypes.c_float] * 5 + [i32, vp, vp]),
        "mvae_sgd_momentum_step": (i32, [vp, vp, vp, ll] + [ctypes.c_float] * 3 + [i32, vp, vp]),
        "mvae_gemm_bf16": (i32, [vp, ll, i32, vp, ll, i32, vp, ll, i32, i32, vp, i32, i32, i32, i32, i32, vp, vp]),
        "mvae_sgemm": (i32, [vp, ll, ll, vp, ll, ll, vp, ll, i32, i32, i32, vp, i32, i32, i32, vp]),
        "mvae_sgemm_tc": (i32, [vp, ll, ll, vp, ll, ll, vp, ll, i32, i32, i32, vp, i32, i32, vp, ctypes.c_size_t, vp, vp]),
        "mvae_sgemm_tc_scratch_bytes": (ctypes.c_size_t, [ll, ll]),
        "mvae_l2_probe": (i32, [vp, ctypes.c_size_t, i32, i32, i32, vp]),
    }
    for name, (res, args) in sigs.items():
        fn = getattr(lib, name)  # AttributeError here == header and library out of sync
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()
EXPORTED = [
    "mvae_strerror", "mvae_last_cuda_error", "mvae_launch_count", "mvae_reset_launch_count",
    "mvae_profile_enable", "mvae_profile_read",
    "mvae_cfgb_workspace_bytes", "mvae_cfgb_elbo_step", "mvae_cfgb_elbo_step_graph_create", "mvae_cfgb_elbo_step_phase",
    "mvae_cfgb_elbo_step_phase_graph_create", "mvae_graph_launch",
    "mvae_graph_num_kernel_nodes", "mvae_graph_destroy", "mvae_cfgb_forward", "mvae_cfgb_backward",
    "mvae_cfgb_decode_greedy", "mvae_onehot_to_ids", "mvae_cfgb_read_error", "mvae_gemm_bf16", "mvae_sgemm", "mvae_sgemm_tc", "mvae_sgemm_tc_scratch_bytes", "mvae_l2_probe",
    "mvae_clip_grad_norm", "mvae_adam_step", "mvae_sgd_momentum_step",
    "mvae_cfga_workspace_bytes", "mvae_cfga_elbo_step", "mvae_cfga_elbo_step_graph_create", "mvae_cfga_forward",
    "mvae_cfga_backward", "mvae_cfga_decode", "mvae_cfga_read_error",
    "mvae_ids_to_text", "mvae_text_to_ids", "mvae_binding_workspace_bytes", "mvae_binding_forward", "mvae_binding_backward",
    "mvae_moses_workspace_bytes", "mvae_moses_step", "mvae_moses_step_ex", "mvae_moses_joint_extra_bytes", "mvae_moses_joint_step",
    "mvae_moses_sample", "mvae_moses_sample_graph_create", "mvae_moses_read_error",
]


def check(rc):
    if rc != 0:
        msg = lib.mvae_strerror(rc).decode()
        if rc == -3:
            msg += ": " + lib.mvae_last_cuda_error().decode()
        raise MvaeError(f"libmvae_b200: {msg} (rc={rc})")
