"""Engine — one libldpc_b200 handle per (parity-check matrix, device), plus the device buffers.

PyTorch is used for exactly three things here: allocating device / pinned memory
(``torch.empty``), naming the current CUDA stream, and (in dist.py) the NCCL process group.
All arithmetic happens in the hand-written kernels behind the C ABI (include/ldpc_b200.h).
"""
import ctypes
import hashlib
import threading

import numpy as np

from . import _lib
from ._lib import LdpcError
from .graph import Tables

_NP2DT = {np.dtype(np.float32): _lib.F32, np.dtype(np.float64): _lib.F64, np.dtype(np.float16): _lib.F16}


def packed_row_bytes(n):
    """Bytes of one bit-packed row of n symbols (ldpc_packed_row_bytes): ceil(n / 8) rounded up to 16."""
    return (((int(n) + 7) // 8) + 15) // 16 * 16


def pack_bits(Y, out=None):
    """Hard bits [B, n] (0/1) -> bit-packed rows uint8 [B, packed_row_bytes(n)] for decode_host(..., packed_in=True).
    Bit v of a row is bit (v & 7) of byte (v >> 3) (numpy.packbits, bitorder="little")."""
    Y = np.asarray(Y)
    B, n = Y.shape
    out = np.zeros((B, packed_row_bytes(n)), np.uint8) if out is None else out
    out[:, :(n + 7) // 8] = np.packbits(Y != 0, axis=1, bitorder="little")
    out[:, (n + 7) // 8:] = 0
    return out


def pack_symbols(Y, out=None):
    """BEC symbols [B, n] in {0, 1, 2} -> two packed planes per row: value plane (== 1), then erasure plane (== 2)."""
    Y = np.asarray(Y)
    B, n = Y.shape
    s = packed_row_bytes(n)
    out = np.zeros((B, 2 * s), np.uint8) if out is None else out
    pack_bits(Y == 1, out[:, :s])
    pack_bits(Y == 2, out[:, s:])
    return out


def unpack_bits(P, n):
    """Inverse of pack_bits: packed rows -> uint8 [B, n]."""
    return np.unpackbits(np.asarray(P)[:, :(n + 7) // 8], axis=1, count=n, bitorder="little")


def unpack_symbols(P, n):
    """Inverse of pack_symbols: two packed planes -> uint8 [B, n] symbols {0, 1, 2}."""
    s = packed_row_bytes(n)
    val, er = unpack_bits(np.asarray(P)[:, :s], n), unpack_bits(np.asarray(P)[:, s:], n)
    return np.where(er != 0, 2, val).astype(np.uint8)


def _torch():
    import torch
    if not torch.cuda.is_available():
        raise LdpcError("no CUDA device: ldpc_decoders_b200 has no CPU fallback")
    return torch


class Engine:
    """Decoder handle for one H on one GPU."""

    def __init__(self, tables, device=None):
        torch = _torch()
        self.lib = _lib.load()
        self.tables = tables
        self.device = torch.cuda.current_device() if device is None else int(device)
        t = tables
        h = ctypes.c_void_p()
        rc = self.lib.ldpc_create(ctypes.byref(h), self.device, t.n, t.m, t.E,
                                  t.chk_ptr.ctypes.data, t.edge_var.ctypes.data,
                                  t.var_ptr.ctypes.data, t.var_edges.ctypes.data)
        if rc != 0:
            raise LdpcError("ldpc_create failed (%d): %s" % (rc, self.lib.ldpc_last_error(None).decode()))
        self.handle = h
        self._ws = None

    def close(self):
        if getattr(self, "handle", None):
            self.lib.ldpc_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ helpers
    @property
    def launch_count(self):
        return int(self.lib.ldpc_launch_count(self.handle))

    @property
    def resident_frames(self):
        """Frames per CTA of the on-chip path for this code; 0 = the code does not fit in shared memory."""
        return int(self.lib.ldpc_resident_frames(self.handle))

    @property
    def resident_kernel(self):
        """Name of the on-chip kernel this code runs on ("resident_vp", "resident_bp") or "" (streaming only)."""
        return self.lib.ldpc_resident_kernel(self.handle).decode()

    def resident_plan(self):
        """Predicted shared-memory wavefronts per iteration of the on-chip path (ldpc_resident_plan), or None."""
        out = (ctypes.c_long * 7)()
        if self.lib.ldpc_resident_plan(self.handle, out) != 0:
            return None
        keys = ("cn_ideal", "cn_file", "cn_plan_natural", "cn_plan", "vn_ideal", "vn_file", "vn_plan")
        return dict(zip(keys, [int(x) for x in out]))

    def profile(self, on):
        """Record CUDA events around every CN / VN sweep launch (see ldpc_profile_enable)."""
        _lib.check(self.handle, self.lib.ldpc_profile_enable(self.handle, 1 if on else 0))

    def profile_read(self):
        """dict(cn_ms, cn_launches, vn_ms, vn_launches) accumulated since the last read; synchronises."""
        cn, vn = ctypes.c_double(), ctypes.c_double()
        ncn, nvn = ctypes.c_ulonglong(), ctypes.c_ulonglong()
        _lib.check(self.handle, self.lib.ldpc_profile_read(self.handle, ctypes.byref(cn), ctypes.byref(ncn),
                                                           ctypes.byref(vn), ctypes.byref(nvn)))
        return dict(cn_ms=cn.value, cn_launches=ncn.value, vn_ms=vn.value, vn_launches=nvn.value)

    def _dev(self):
        return _torch().device("cuda", self.device)

    def workspace(self, algo, dtype, B, flags=0):
        """A cached torch uint8 buffer of at least ldpc_workspace_bytes (grow-only)."""
        torch = _torch()
        need = int(self.lib.ldpc_workspace_bytes(self.handle, algo, dtype, int(B), flags))
        if need == 0:
            raise LdpcError("ldpc_workspace_bytes: bad arguments")
        if self._ws is None or self._ws.numel() < need:
            self._ws = None
            self._ws = torch.empty(need, dtype=torch.uint8, device=self._dev())
        return self._ws

    @staticmethod
    def _stream_ptr(stream):
        torch = _torch()
        s = torch.cuda.current_stream() if stream is None else stream
        return ctypes.c_void_p(s.cuda_stream)

    # ------------------------------------------------------------------ device-resident decode
    def decode_device(self, algo, inp, y_hard=None, max_iter=10, iter_cap=0, want_marg=False,
                      want_reason=True, flags=0, stream=None, out=None):
        """Decode a batch that already lives on the GPU.

        inp: torch CUDA tensor [B, n]; float32/float64 priors for MSA/SPA (dtype selects the
        arithmetic, like the reference's priors.dtype), uint8 symbols for BEC.
        Returns dict(x_hat uint8 [B,n], iters int32 [B], reason uint8 [B] | None, marg | None).
        Asynchronous on the current stream.
        """
        torch = _torch()
        t = self.tables
        if inp.dim() != 2 or inp.shape[1] != t.n or not inp.is_cuda or not inp.is_contiguous():
            raise ValueError("input must be a contiguous CUDA tensor [B, n]")
        B = int(inp.shape[0])
        if algo == _lib.BEC:
            if inp.dtype != torch.uint8:
                raise TypeError("BEC input must be uint8 symbols {0,1,2}")
            dtype = _lib.F32
        else:
            if inp.dtype == torch.float32:
                dtype = _lib.F32
            elif inp.dtype == torch.float64:
                dtype = _lib.F64
            else:
                raise TypeError("priors must be float32 or float64")
        if y_hard is not None and (y_hard.dtype != torch.uint8 or tuple(y_hard.shape) != (B, t.n)
                                   or not y_hard.is_contiguous()):
            raise ValueError("y_hard must be a contiguous uint8 CUDA tensor [B, n]")
        dev = self._dev()
        if out is None:
            out = {}
        for key, shape in (("x_hat", (B, t.n)), ("iters", (B,))):        # buffers of another batch size are not reused
            if out.get(key) is not None and tuple(out[key].shape) != shape:
                out[key] = None
        x_hat = out.get("x_hat")
        if x_hat is None:
            x_hat = torch.empty((B, t.n), dtype=torch.uint8, device=dev)
        iters = out.get("iters")
        if iters is None:
            iters = torch.empty(B, dtype=torch.int32, device=dev)
        reason = torch.empty(B, dtype=torch.uint8, device=dev) if want_reason else None
        marg = torch.empty((B, t.n), dtype=inp.dtype, device=dev) if (want_marg and algo != _lib.BEC) else None
        ws = self.workspace(algo, dtype, B, flags)
        rc = self.lib.ldpc_decode(self.handle, algo, dtype, inp.data_ptr(),
                                  None if y_hard is None else y_hard.data_ptr(), B,
                                  int(max_iter), int(iter_cap), x_hat.data_ptr(), iters.data_ptr(),
                                  None if reason is None else reason.data_ptr(),
                                  None if marg is None else marg.data_ptr(),
                                  ws.data_ptr(), ws.numel(), flags, self._stream_ptr(stream))
        _lib.check(self.handle, rc)
        return dict(x_hat=x_hat, iters=iters, reason=reason, marg=marg)

    def decode_device_channel(self, channel, algo, dtype, param, y, max_iter=10, iter_cap=0, want_reason=True,
                              flags=0, stream=None, out=None):
        """Like decode_device, but takes the RECEIVED block y [B,n] (CUDA tensor: uint8 for BSC/BEC, float32 or
        float64 for BIAWGN) and applies the channel's LLR map inside the load kernel (ldpc_decode_channel)."""
        torch = _torch()
        t = self.tables
        if y.dim() != 2 or y.shape[1] != t.n or not y.is_cuda or not y.is_contiguous():
            raise ValueError("y must be a contiguous CUDA tensor [B, n]")
        B = int(y.shape[0])
        if channel in (_lib.CH_BSC, _lib.CH_BEC):
            if y.dtype != torch.uint8:
                raise TypeError("BSC/BEC input must be uint8")
            y_dtype = _lib.F32
        else:
            y_dtype = {torch.float32: _lib.F32, torch.float64: _lib.F64, torch.float16: _lib.F16}[y.dtype]
        dev = self._dev()
        out = {} if out is None else out
        for key, shape in (("x_hat", (B, t.n)), ("iters", (B,)), ("reason", (B,))):     # buffers of another batch size
            if out.get(key) is not None and tuple(out[key].shape) != shape:
                out[key] = None
        x_hat = out.get("x_hat")
        if x_hat is None:
            x_hat = torch.empty((B, t.n), dtype=torch.uint8, device=dev)
        iters = out.get("iters")
        if iters is None:
            iters = torch.empty(B, dtype=torch.int32, device=dev)
        reason = out.get("reason")
        if reason is None and want_reason:
            reason = torch.empty(B, dtype=torch.uint8, device=dev)
        ws = self.workspace(algo, dtype, B, flags)
        rc = self.lib.ldpc_decode_channel(self.handle, channel, algo, dtype, float(param), y.data_ptr(), y_dtype, B,
                                          int(max_iter), int(iter_cap), x_hat.data_ptr(), iters.data_ptr(),
                                          None if reason is None else reason.data_ptr(), None,
                                          ws.data_ptr(), ws.numel(), flags, self._stream_ptr(stream))
        _lib.check(self.handle, rc)
        return dict(x_hat=x_hat, iters=iters, reason=reason, marg=None)

    # ------------------------------------------------------------------ on-device Monte-Carlo round
    def channel_generate(self, channel, param, B, seed, frame0=0, x=None, stream=None, out=None):
        """Received block for B frames of the word x (CUDA uint8 [n], None = all-zero) drawn on the GPU
        (ldpc_channel_generate): float32 [B,n] for BIAWGN (param = noise_var), uint8 for BSC / BEC (param = p)."""
        torch = _torch()
        t = self.tables
        dt = torch.float32 if channel == _lib.CH_BIAWGN else torch.uint8
        if out is not None and (tuple(out.shape) != (int(B), t.n) or out.dtype != dt or not out.is_contiguous()):
            out = None                         # a cached block of another batch size / channel: never write past it
        y = out if out is not None else torch.empty((int(B), t.n), dtype=dt, device=self._dev())
        rc = self.lib.ldpc_channel_generate(self.handle, channel, float(param), None if x is None else x.data_ptr(),
                                            int(seed) & (2 ** 64 - 1), int(frame0), int(B), y.data_ptr(),
                                            self._stream_ptr(stream))
        _lib.check(self.handle, rc)
        return y

    def count_errors(self, x_hat, x=None, stream=None):
        """int32 [B]: bit errors of every decoded frame against the word x (None = all-zero)."""
        torch = _torch()
        errs = torch.empty(int(x_hat.shape[0]), dtype=torch.int32, device=x_hat.device)
        rc = self.lib.ldpc_count_errors(self.handle, x_hat.data_ptr(), None if x is None else x.data_ptr(),
                                        int(x_hat.shape[0]), errs.data_ptr(), self._stream_ptr(stream))
        _lib.check(self.handle, rc)
        return errs

    def simulate(self, channel, algo, dtype, param, B, seed, frame0=0, x=None, max_iter=10, iter_cap=0, flags=0,
                 stream=None, bufs=None):
        """One Monte-Carlo round entirely on the GPU: draw B received frames (global indices frame0 ..), decode,
        count bit errors.  `param` is the reference's channel parameter (p, or the SNR-derived noise_var for
        BIAWGN; the BSC decoder gets llr = log(1-p) - log(p)).  Returns dict(bit_errs int32 [B], iters int32 [B],
        reason uint8 [B], x_hat uint8 [B,n]) of CUDA tensors; only the two small vectors need to travel to the host."""
        bufs = {} if bufs is None else bufs
        y = self.channel_generate(channel, param, B, seed, frame0, x, stream, out=bufs.get("y"))
        bufs["y"] = y
        dec_param = float(np.log(1 - param) - np.log(param)) if channel == _lib.CH_BSC else param
        out = self.decode_device_channel(channel, _lib.BEC if channel == _lib.CH_BEC else algo, dtype, dec_param, y,
                                         max_iter=max_iter, iter_cap=iter_cap, flags=flags, stream=stream,
                                         out=bufs.get("out"))
        bufs["out"] = out
        out["bit_errs"] = self.count_errors(out["x_hat"], x, stream)
        return out

    def new_counters(self, nhist):
        """Zeroed device int64 [4 + nhist] Monte-Carlo counters for mc_round: tot, wec, bec, sum of iters, histogram."""
        torch = _torch()
        return torch.zeros(4 + int(nhist), dtype=torch.int64, device=self._dev())

    def mc_round(self, channel, algo, dtype, param, B, seed, frame0, counters, nhist, x=None, max_iter=10, iter_cap=0,
                 flags=0, stream=None):
        """One Monte-Carlo round that never leaves the GPU (ldpc_mc_round): draw frames frame0 .. frame0 + B - 1 of the
        word x, decode, add (tot, wec, bec, sum iters, iteration histogram) to the device tensor `counters`
        (new_counters(nhist)); asynchronous on the current stream.  `param` as in simulate()."""
        torch = _torch()
        if counters.dtype != torch.int64 or counters.numel() < 4 + int(nhist) or not counters.is_cuda:
            raise ValueError("counters must be a CUDA int64 tensor of 4 + nhist elements")
        dec_algo = _lib.BEC if channel == _lib.CH_BEC else algo
        need = int(self.lib.ldpc_mc_scratch_bytes(self.handle, channel, dec_algo, dtype, int(B)))
        if need == 0:
            raise LdpcError("ldpc_mc_scratch_bytes: bad arguments")
        if getattr(self, "_mc_scratch", None) is None or self._mc_scratch.numel() < need:
            self._mc_scratch = None
            self._mc_scratch = torch.empty(need, dtype=torch.uint8, device=self._dev())
        dec_param = float(np.log(1 - param) - np.log(param)) if channel == _lib.CH_BSC else float(param)
        rc = self.lib.ldpc_mc_round(self.handle, channel, dec_algo, dtype, float(param), dec_param,
                                    None if x is None else x.data_ptr(), int(seed) & (2 ** 64 - 1), int(frame0), int(B),
                                    int(max_iter), int(iter_cap), counters.data_ptr(), int(nhist),
                                    self._mc_scratch.data_ptr(), self._mc_scratch.numel(), flags, self._stream_ptr(stream))
        _lib.check(self.handle, rc)

    # ------------------------------------------------------------------ host-buffer decode (e2e path)
    def decode_host(self, channel, algo, dtype, param, y, max_iter=10, iter_cap=0, chunk=0, flags=0,
                    x_hat=None, iters=None, reason=None, wait=True, packed_in=False, packed_out=False):
        """Decode a batch held in host memory (numpy); H2D / decode / D2H are pipelined in the library.

        y [B, n]: uint8 for BSC/BEC, float32 / float64 / float16 for BIAWGN / PRIORS.  Pinned arrays
        (``pinned_empty``) make the copies asynchronous.  Returns numpy (x_hat, iters, reason).
        packed_in (BSC / BEC): y is bit-packed rows (pack_bits / pack_symbols) — 1 bit per hard bit on PCIe instead of
        a byte; packed_out: x_hat comes back bit-packed the same way (unpack_bits / unpack_symbols).
        wait=False (a stream of batches): returns once the work is enqueued, so the next call overlaps this one's
        tail; y and the output arrays must be pinned, caller-provided and left alone until ``host_sync()``.
        """
        if not wait:
            if x_hat is None or iters is None or reason is None:
                raise ValueError("wait=False needs caller-provided (pinned) output arrays")
            flags |= _lib.HOST_ASYNC
        t = self.tables
        y = np.ascontiguousarray(y)
        planes = 2 if (channel == _lib.CH_BEC or algo == _lib.BEC) else 1
        prow = planes * packed_row_bytes(t.n)
        if packed_in:
            if channel not in (_lib.CH_BSC, _lib.CH_BEC):
                raise ValueError("packed_in is for BSC / BEC symbol input")
            if y.ndim != 2 or y.shape[1] != prow or y.dtype != np.uint8:
                raise ValueError("packed y must be uint8 [B, %d]" % prow)
            flags |= _lib.IN_PACKED
        elif y.ndim != 2 or y.shape[1] != t.n:
            raise ValueError("y must be [B, n]")
        B = y.shape[0]
        if channel in (_lib.CH_BSC, _lib.CH_BEC):
            if y.dtype != np.uint8:
                raise TypeError("BSC/BEC input must be uint8")
            y_dtype = _lib.F32
        else:
            if y.dtype not in _NP2DT:
                raise TypeError("input must be float16, float32 or float64")
            y_dtype = _NP2DT[y.dtype]
        if B == 0:                                          # an empty batch decodes to empty results (the C ABI wants B > 0)
            return (np.empty((0, prow if packed_out else t.n), np.uint8), np.empty(0, np.int32), np.empty(0, np.uint8))
        if packed_out:
            flags |= _lib.OUT_PACKED
        xshape = (B, prow) if packed_out else (B, t.n)
        if x_hat is None:
            x_hat = np.empty(xshape, np.uint8)
        elif tuple(x_hat.shape) != xshape or x_hat.dtype != np.uint8 or not x_hat.flags.c_contiguous:
            raise ValueError("x_hat must be a contiguous uint8 array of shape %r" % (xshape,))
        if iters is None:
            iters = np.empty(B, np.int32)
        if reason is None:
            reason = np.empty(B, np.uint8)
        if iters.shape != (B,) or reason.shape != (B,):
            raise ValueError("iters / reason must have one element per frame")
        rc = self.lib.ldpc_decode_host(self.handle, channel, algo, dtype, float(param), y.ctypes.data, y_dtype,
                                       B, int(max_iter), int(iter_cap), x_hat.ctypes.data, iters.ctypes.data,
                                       reason.ctypes.data, int(chunk), flags)
        _lib.check(self.handle, rc)
        return x_hat, iters, reason

    def host_sync(self):
        """Wait for every decode_host(..., wait=False) enqueued so far (ldpc_host_sync)."""
        _lib.check(self.handle, self.lib.ldpc_host_sync(self.handle))

    # ------------------------------------------------------------------ front ends and the isolated sweep
    def llr_bsc(self, p_llr, y, dtype, stream=None):
        torch = _torch()
        out = torch.empty(y.shape, dtype=torch.float32 if dtype == _lib.F32 else torch.float64, device=y.device)
        rc = self.lib.ldpc_llr_bsc(self.handle, dtype, float(p_llr), y.data_ptr(), out.data_ptr(), y.numel(),
                                   self._stream_ptr(stream))
        _lib.check(self.handle, rc)
        return out

    def llr_biawgn(self, noise_var, y, dtype, stream=None):
        torch = _torch()
        y_dtype = _lib.F32 if y.dtype == torch.float32 else _lib.F64
        out = torch.empty(y.shape, dtype=torch.float32 if dtype == _lib.F32 else torch.float64, device=y.device)
        rc = self.lib.ldpc_llr_biawgn(self.handle, y_dtype, dtype, float(noise_var), y.data_ptr(), out.data_ptr(),
                                      y.numel(), self._stream_ptr(stream))
        _lib.check(self.handle, rc)
        return out

    def debug_step(self, algo, which, msg_in, prior=None, stream=None):
        """One CN (which=0) or VN (which=1) sweep on messages [B, E] in np.where(H) edge order."""
        torch = _torch()
        t = self.tables
        if msg_in.dim() != 2 or msg_in.shape[1] != t.E or not msg_in.is_cuda or not msg_in.is_contiguous():
            raise ValueError("msg_in must be a contiguous CUDA tensor [B, E]")
        if msg_in.dtype not in (torch.float32, torch.float64):
            raise TypeError("messages must be float32 or float64")
        if prior is not None and (tuple(prior.shape) != (msg_in.shape[0], t.n) or prior.dtype != msg_in.dtype or not prior.is_contiguous()):
            raise ValueError("prior must be a contiguous [B, n] tensor of the message dtype")
        B = int(msg_in.shape[0])
        dtype = _lib.F32 if msg_in.dtype == torch.float32 else _lib.F64
        msg_out = torch.empty_like(msg_in)
        marg = torch.empty((B, t.n), dtype=msg_in.dtype, device=msg_in.device) if which == 1 else None
        ws = self.workspace(algo, dtype, B)
        rc = self.lib.ldpc_debug_step(self.handle, algo, dtype, which, B,
                                      None if prior is None else prior.data_ptr(), msg_in.data_ptr(),
                                      msg_out.data_ptr(), None if marg is None else marg.data_ptr(),
                                      ws.data_ptr(), ws.numel(), self._stream_ptr(stream))
        _lib.check(self.handle, rc)
        return msg_out, marg


def pinned_empty(shape, dtype):
    """A numpy array backed by pinned (page-locked) host memory, for asynchronous copies."""
    torch = _torch()
    tdt = {np.dtype(np.uint8): torch.uint8, np.dtype(np.int32): torch.int32,
           np.dtype(np.float32): torch.float32, np.dtype(np.float64): torch.float64,
           np.dtype(np.float16): torch.float16, np.dtype(np.int64): torch.int64}[np.dtype(dtype)]
    t = torch.empty(tuple(shape) if not np.isscalar(shape) else (shape,), dtype=tdt, pin_memory=True)
    return t.numpy()          # the array keeps the tensor (and its pinned allocation) alive


_engines = {}
_engines_lock = threading.Lock()


def tables_of(code_or_mtx):
    """Tables from a reference-style Code object (``.parity_mtx``), a dense H, or ready-made Tables."""
    if isinstance(code_or_mtx, Tables):
        return code_or_mtx
    tab = getattr(code_or_mtx, "tables", None)
    if isinstance(tab, Tables):
        return tab
    H = getattr(code_or_mtx, "parity_mtx", code_or_mtx)
    return Tables.from_dense(H)


def engine_for(tables, device=None):
    """Engines are cached per (edge list, device): src/main.py builds a new decoder per channel parameter."""
    torch = _torch()
    dev = torch.cuda.current_device() if device is None else int(device)
    key = (tables.m, tables.n, hashlib.sha1(tables.edge_var.tobytes() + tables.chk_ptr.tobytes()).hexdigest(), dev)
    with _engines_lock:
        eng = _engines.get(key)
        if eng is None:
            eng = Engine(tables, dev)
            _engines[key] = eng
        return eng
