"""ctypes binding of libldpc_b200.so (include/ldpc_b200.h).  There is no fallback: if the library
is missing or the CUDA device is absent, every entry point raises."""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libldpc_b200.so")

MSA, SPA, BEC = 0, 1, 2
F32, F64, F16 = 0, 1, 2
CH_PRIORS, CH_BSC, CH_BIAWGN, CH_BEC = 0, 1, 2, 3
PATH_AUTO, PATH_STREAMING, PATH_RESIDENT = 0, 1, 2
SPA_ROBUST = 4
CN_REGISTER = 8
HOST_ASYNC = 16
OUT_PACKED = 32
IN_PACKED = 64
REASONS = {0: "decoded", 1: "maximum", 2: "stopping", 4: "cap"}

# every symbol include/ldpc_b200.h declares
SYMBOLS = ("ldpc_abi_version", "ldpc_create", "ldpc_destroy", "ldpc_last_error", "ldpc_workspace_bytes",
           "ldpc_decode", "ldpc_decode_channel", "ldpc_llr_bsc", "ldpc_llr_biawgn", "ldpc_debug_step", "ldpc_decode_host", "ldpc_host_sync",
           "ldpc_launch_count", "ldpc_profile_enable", "ldpc_profile_read", "ldpc_resident_frames", "ldpc_resident_plan", "ldpc_resident_kernel", "ldpc_channel_generate", "ldpc_count_errors",
           "ldpc_packed_row_bytes", "ldpc_count_accumulate", "ldpc_mc_scratch_bytes", "ldpc_mc_round")


class LdpcError(RuntimeError):
    pass


_lib = None


def load():
    """Load the CUDA library; raises LdpcError (never falls back to a CPU path)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LdpcError("libldpc_b200.so is not built: run `python -m ldpc_decoders_b200.build` "
                        "(or __graft_entry__.build()); there is no CPU fallback")
    L = ctypes.CDLL(LIB_PATH)
    vp, i32, u32, sz, dbl = ctypes.c_void_p, ctypes.c_int, ctypes.c_uint, ctypes.c_size_t, ctypes.c_double
    L.ldpc_abi_version.restype = i32
    L.ldpc_abi_version.argtypes = []
    L.ldpc_create.restype = i32
    L.ldpc_create.argtypes = [ctypes.POINTER(vp), i32, i32, i32, i32, vp, vp, vp, vp]
    L.ldpc_destroy.restype = None
    L.ldpc_destroy.argtypes = [vp]
    L.ldpc_last_error.restype = ctypes.c_char_p
    L.ldpc_last_error.argtypes = [vp]
    L.ldpc_workspace_bytes.restype = sz
    L.ldpc_workspace_bytes.argtypes = [vp, i32, i32, i32, u32]
    L.ldpc_decode.restype = i32
    L.ldpc_decode.argtypes = [vp, i32, i32, vp, vp, i32, i32, i32, vp, vp, vp, vp, vp, sz, u32, vp]
    L.ldpc_decode_channel.restype = i32
    L.ldpc_decode_channel.argtypes = [vp, i32, i32, i32, dbl, vp, i32, i32, i32, i32, vp, vp, vp, vp, vp, sz, u32, vp]
    L.ldpc_llr_bsc.restype = i32
    L.ldpc_llr_bsc.argtypes = [vp, i32, dbl, vp, vp, sz, vp]
    L.ldpc_llr_biawgn.restype = i32
    L.ldpc_llr_biawgn.argtypes = [vp, i32, i32, dbl, vp, vp, sz, vp]
    L.ldpc_debug_step.restype = i32
    L.ldpc_debug_step.argtypes = [vp, i32, i32, i32, i32, vp, vp, vp, vp, vp, sz, vp]
    L.ldpc_decode_host.restype = i32
    L.ldpc_decode_host.argtypes = [vp, i32, i32, i32, dbl, vp, i32, i32, i32, i32, vp, vp, vp, i32, u32]
    L.ldpc_profile_enable.restype = i32
    L.ldpc_profile_enable.argtypes = [vp, i32]
    L.ldpc_profile_read.restype = i32
    L.ldpc_profile_read.argtypes = [vp, ctypes.POINTER(dbl), ctypes.POINTER(ctypes.c_ulonglong),
                                    ctypes.POINTER(dbl), ctypes.POINTER(ctypes.c_ulonglong)]
    L.ldpc_resident_frames.restype = i32
    L.ldpc_resident_frames.argtypes = [vp]
    L.ldpc_channel_generate.restype = i32
    L.ldpc_channel_generate.argtypes = [vp, i32, dbl, vp, ctypes.c_ulonglong, ctypes.c_ulonglong, i32, vp, vp]
    L.ldpc_count_errors.restype = i32
    L.ldpc_count_errors.argtypes = [vp, vp, vp, i32, vp, vp]
    L.ldpc_packed_row_bytes.restype = sz
    L.ldpc_packed_row_bytes.argtypes = [i32]
    L.ldpc_count_accumulate.restype = i32
    L.ldpc_count_accumulate.argtypes = [vp, vp, vp, vp, i32, vp, vp, i32, vp]
    L.ldpc_mc_scratch_bytes.restype = sz
    L.ldpc_mc_scratch_bytes.argtypes = [vp, i32, i32, i32, i32]
    L.ldpc_mc_round.restype = i32
    L.ldpc_mc_round.argtypes = [vp, i32, i32, i32, dbl, dbl, vp, ctypes.c_ulonglong, ctypes.c_ulonglong, i32, i32, i32,
                                vp, i32, vp, sz, u32, vp]
    L.ldpc_host_sync.restype = i32
    L.ldpc_host_sync.argtypes = [vp]
    L.ldpc_resident_kernel.restype = ctypes.c_char_p
    L.ldpc_resident_kernel.argtypes = [vp]
    L.ldpc_resident_plan.restype = i32
    L.ldpc_resident_plan.argtypes = [vp, ctypes.POINTER(ctypes.c_long)]
    L.ldpc_launch_count.restype = ctypes.c_ulonglong
    L.ldpc_launch_count.argtypes = [vp]
    if L.ldpc_abi_version() != 2:
        raise LdpcError("libldpc_b200.so ABI version mismatch")
    _lib = L
    return L


def check(handle, rc):
    if rc != 0:
        msg = load().ldpc_last_error(handle)
        raise LdpcError("libldpc_b200 error %d: %s" % (rc, (msg or b"").decode(errors="replace")))
