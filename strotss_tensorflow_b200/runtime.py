"""Handle management: one library handle per (owner, CUDA device); torch supplies device memory and
the current stream, nothing else."""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import torch

from . import _lib


def _ptr(t: Optional[torch.Tensor]):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _stream(device: torch.device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _check_features(name: str, t: torch.Tensor) -> torch.Tensor:
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor")
    if not t.is_cuda:
        raise RuntimeError(f"{name} is on {t.device}: the STROTSS loss path runs on a B200 only (no CPU fallback)")
    if t.dtype != torch.float32:
        raise TypeError(f"{name} must be float32 (the reference computes in fp32), got {t.dtype}")
    return t


def _rows(t: torch.Tensor) -> torch.Tensor:
    """Row-major (n, C) view the C ABI can read in place: unit column stride, any row stride >= C (a column slice of a wider
    matrix is passed with its own `ld` instead of being copied)."""
    if t.dim() == 2 and t.stride(1) == 1 and t.stride(0) >= t.shape[1]:
        return t
    return t.contiguous()


def reshape_2d(x: torch.Tensor, channel_axis: int = -1) -> torch.Tensor:
    """nn/losses.py:31-36: squeeze, then reshape to (-1, C).  (The rank test at :32 never fires.)"""
    x = torch.squeeze(x)
    if x.dim() == 0:
        x = x.reshape(1, 1)
    return x.reshape(-1, x.shape[channel_axis])


class Handle:
    """Owns one strotss_handle (device workspace + cached style target)."""

    def __init__(self, device: torch.device):
        self.lib = _lib.load()
        if device.type != "cuda":
            raise RuntimeError("strotss_tensorflow_b200 needs a CUDA (sm_100) device; there is no CPU path")
        self.device = device
        idx = device.index if device.index is not None else torch.cuda.current_device()
        self.index = idx
        h = C.c_void_p()
        code = self.lib.strotss_create(idx, C.byref(h))
        self._h = h
        if code != 0:
            msg = self.lib.strotss_last_error(h).decode() if h else "allocation failed"
            if h:
                self.lib.strotss_destroy(h)
            self._h = None
            raise _lib.StrotssError(f"strotss_create(device={idx}) failed (code {code}): {msg}")
        self.style_shape = None
        self.rank, self.world = 0, 1

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                torch.cuda.synchronize(self.device)
            except Exception:
                pass
            self.lib.strotss_destroy(h)
            self._h = None

    def _ck(self, code, what):
        _lib.check(self.lib, self._h, code, what)

    @property
    def launch_count(self) -> int:
        return int(self.lib.strotss_launch_count(self._h))

    @property
    def workspace_bytes(self) -> int:
        return int(self.lib.strotss_workspace_bytes(self._h))

    # ---- multi-GPU ------------------------------------------------------------------
    def comm_init(self, rank: int, world: int, unique_id):
        """Attach this handle to an NCCL communicator (see distributed.attach)."""
        self._ck(self.lib.strotss_comm_init(self._h, int(rank), int(world), unique_id), "strotss_comm_init")
        self.rank, self.world = int(rank), int(world)

    def shard_rows(self, N: int):
        r0, r1 = C.c_int(), C.c_int()
        self._ck(self.lib.strotss_shard_rows(self._h, int(N), C.byref(r0), C.byref(r1)), "strotss_shard_rows")
        return r0.value, r1.value

    def comm_transport(self) -> int:
        """1: sign blocks through CUDA-IPC peer windows, -1: fp32 products through ncclSend/ncclRecv, 0: undecided / single GPU."""
        return int(self.lib.strotss_comm_transport(self._h))

    def collectives_note(self, M: int, D: int) -> str:
        t = self.comm_transport()
        small = (f"two small collectives per evaluation: {{sum over N floats (r), max over 2x{M} packed u64 minima}} and a sum of "
                 f"{16 + D} floats")
        if t == 1:
            return (small + " -- one-shot allreduces through CUDA-IPC peer windows (peer stores + epoch flags over NVLink; "
                    "STROTSS_PEER_AR=0: NCCL); mirrored self-similarity tiles: bf16 sign blocks by copy engine into the owners' "
                    "windows; covariance forward: partial Gram tiles by epilogue peer stores, sign tiles back to every rank")
        if t == -1:
            return small + " (NCCL); mirrored self-similarity tiles: fp32 products by ncclSend/ncclRecv (peer windows unavailable)"
        return small + " (NCCL)"

    def profile_enable(self, on: bool = True):
        self._ck(self.lib.strotss_profile_enable(self._h, 1 if on else 0), "strotss_profile_enable")

    def profile_read(self):
        """-> {phase: (total_ms, launches)} since the last read (synchronises on the recorded events)."""
        n = self.lib.strotss_profile_num_phases()
        ms = (C.c_double * n)()
        cnt = (C.c_longlong * n)()
        self._ck(self.lib.strotss_profile_read(self._h, ms, cnt), "strotss_profile_read")
        return {self.lib.strotss_profile_phase_name(i).decode(): (ms[i], int(cnt[i])) for i in range(n)}

    # ---- fused path -----------------------------------------------------------------
    def set_style_target(self, style: torch.Tensor):
        style = _rows(_check_features("style target", reshape_2d(style)))
        M, D = style.shape
        self._ck(self.lib.strotss_set_style_target(self._h, _ptr(style), M, D, style.stride(0), _stream(style.device)),
                 "strotss_set_style_target")
        self.style_shape = (M, D)

    def eval(self, pred: torch.Tensor, content: torch.Tensor, alpha: float, want_grad: bool = True, want_argmin: bool = False):
        pred = _rows(_check_features("prediction", reshape_2d(pred)))
        content = _rows(_check_features("content", reshape_2d(content)))
        if pred.shape != content.shape:
            raise ValueError(f"prediction {tuple(pred.shape)} and content {tuple(content.shape)} must have the same shape")
        N, D = pred.shape
        if self.style_shape is None or self.style_shape[1] != D:
            raise ValueError("style target not set or feature width differs from the prediction's")
        scalars = torch.empty(_lib.NUM_SCALARS, device=pred.device, dtype=torch.float32)
        grad = torch.empty(N, D, device=pred.device, dtype=torch.float32) if want_grad else None
        ra = torch.empty(self.style_shape[0], device=pred.device, dtype=torch.int32) if want_argmin else None
        ca = torch.empty(N, device=pred.device, dtype=torch.int32) if want_argmin else None
        self._ck(self.lib.strotss_eval(self._h, _ptr(pred), pred.stride(0), _ptr(content), content.stride(0), N, float(alpha),
                                       _ptr(scalars), _ptr(grad), D, _ptr(ra), _ptr(ca), _stream(pred.device)),
                 "strotss_eval")
        return scalars, grad, ra, ca

    # ---- masked (region-guided) transfer: R ragged problems per evaluation ---------------------
    def set_style_targets_grouped(self, styles):
        """styles: list of (M_r, D) tensors, one per region (run_strotss.py:99-101)."""
        styles = [_check_features("style target", reshape_2d(s)) for s in styles]
        if not styles:
            raise ValueError("at least one region is required")
        cat = torch.cat(styles, dim=0).contiguous()
        offs = [0]
        for s in styles:
            offs.append(offs[-1] + int(s.shape[0]))
        R, D = len(styles), int(cat.shape[1])
        arr = (C.c_int * (R + 1))(*offs)
        self._ck(self.lib.strotss_set_style_targets_grouped(self._h, _ptr(cat), cat.stride(0), arr, R, D, _stream(cat.device)),
                 "strotss_set_style_targets_grouped")
        self.group_shape = (R, D)

    def eval_grouped(self, preds, contents, alpha: float, want_grad: bool = True):
        """preds / contents: lists of (N_r, D) tensors (run_strotss.py:114-121).  Returns (mean scalars [16],
        per-region scalars [R, 16], list of per-region gradients or None)."""
        if getattr(self, "group_shape", None) is None or len(preds) != self.group_shape[0] or len(contents) != len(preds):
            raise ValueError("number of regions differs from set_style_targets_grouped")
        preds = [_check_features("prediction", reshape_2d(p)) for p in preds]
        contents = [_check_features("content", reshape_2d(c)) for c in contents]
        offs = [0]
        for p, c in zip(preds, contents):
            if p.shape != c.shape:
                raise ValueError(f"prediction {tuple(p.shape)} and content {tuple(c.shape)} must have the same shape")
            offs.append(offs[-1] + int(p.shape[0]))
        pcat = torch.cat(preds, dim=0).contiguous()
        ccat = torch.cat(contents, dim=0).contiguous()
        R, D = self.group_shape
        if pcat.shape[1] != D:
            raise ValueError("feature width differs from the style targets'")
        arr = (C.c_int * (R + 1))(*offs)
        scalars = torch.empty(_lib.NUM_SCALARS, device=pcat.device, dtype=torch.float32)
        region = torch.empty(R, _lib.NUM_SCALARS, device=pcat.device, dtype=torch.float32)
        grad = torch.empty_like(pcat) if want_grad else None
        self._ck(self.lib.strotss_eval_grouped(self._h, _ptr(pcat), pcat.stride(0), _ptr(ccat), ccat.stride(0), arr, R, float(alpha),
                                               _ptr(scalars), _ptr(region), _ptr(grad), D, _stream(pcat.device)),
                 "strotss_eval_grouped")
        grads = list(torch.split(grad, [b - a for a, b in zip(offs[:-1], offs[1:])], dim=0)) if want_grad else None
        return scalars, region, grads

    def eval_host(self, pred_host: torch.Tensor, content_host: torch.Tensor, alpha: float, grad_host: Optional[torch.Tensor],
                  scalars_host: torch.Tensor):
        """Host-buffer evaluation (bench e2e): tensors are CPU float32, ideally pinned."""
        N, D = pred_host.shape
        dev = self.device
        self._ck(self.lib.strotss_eval_host(self._h, _ptr(pred_host), _ptr(content_host), N, float(alpha), _ptr(scalars_host),
                                            _ptr(grad_host), _stream(dev)), "strotss_eval_host")

    def eval_host_submit(self, pred_host, content_host, alpha: float, grad_host, scalars_host) -> int:
        """Pipelined host-buffer evaluation: returns a ticket; results are valid after eval_host_wait(ticket)."""
        N, D = pred_host.shape
        t = C.c_longlong(-1)
        self._ck(self.lib.strotss_eval_host_submit(self._h, _ptr(pred_host), _ptr(content_host), N, float(alpha), _ptr(scalars_host),
                                                   _ptr(grad_host), C.byref(t)), "strotss_eval_host_submit")
        return int(t.value)

    def eval_host_wait(self, ticket: int):
        self._ck(self.lib.strotss_eval_host_wait(self._h, int(ticket)), "strotss_eval_host_wait")

    def style_loss(self, pred: torch.Tensor, alpha: float, want_grad: bool = True):
        pred = _check_features("prediction", reshape_2d(pred)).contiguous()
        N, D = pred.shape
        if self.style_shape is None or self.style_shape[1] != D:
            raise ValueError("style target not set or feature width differs from the prediction's")
        scalars = torch.empty(_lib.NUM_SCALARS, device=pred.device, dtype=torch.float32)
        grad = torch.empty_like(pred) if want_grad else None
        self._ck(self.lib.strotss_style_loss(self._h, _ptr(pred), pred.stride(0), N, float(alpha), _ptr(scalars), _ptr(grad), D,
                                             _stream(pred.device)), "strotss_style_loss")
        return scalars, grad

    # ---- per-function entry points ----------------------------------------------------
    def relaxed_emd(self, x, y, distance: str, want_grad: bool, want_argmin: bool = False):
        if distance not in _lib.DIST_CODES:
            raise KeyError(distance)                      # nn/losses.py:74
        x = _check_features("x", x).contiguous()
        y = _check_features("y", y).contiguous()
        if x.shape[1] != y.shape[1]:
            raise ValueError("x and y must have the same number of channels")
        M, D = x.shape
        N = y.shape[0]
        out = torch.empty(4, device=y.device, dtype=torch.float32)
        grad = torch.empty_like(y) if want_grad else None
        ra = torch.empty(M, device=y.device, dtype=torch.int32) if want_argmin else None
        ca = torch.empty(N, device=y.device, dtype=torch.int32) if want_argmin else None
        self._ck(self.lib.strotss_relaxed_emd(self._h, _ptr(x), x.stride(0), M, _ptr(y), y.stride(0), N, D,
                                              _lib.DIST_CODES[distance], _ptr(out), _ptr(grad), D, _ptr(ra), _ptr(ca),
                                              _stream(y.device)), "strotss_relaxed_emd")
        return out, grad, ra, ca

    def moment_matching(self, x, y, want_grad: bool):
        x = _check_features("x", x).contiguous()
        y = _check_features("y", y).contiguous()
        if x.shape[1] != y.shape[1]:
            raise ValueError("x and y must have the same number of channels")
        M, D = x.shape
        N = y.shape[0]
        out = torch.empty(3, device=y.device, dtype=torch.float32)
        grad = torch.empty_like(y) if want_grad else None
        self._ck(self.lib.strotss_moment_matching(self._h, _ptr(x), x.stride(0), M, _ptr(y), y.stride(0), N, D, _ptr(out),
                                                  _ptr(grad), D, _stream(y.device)), "strotss_moment_matching")
        return out, grad

    def self_similarity(self, x, y, want_grad: bool):
        x = _check_features("x", x).contiguous()
        y = _check_features("y", y).contiguous()
        if x.shape != y.shape:
            raise ValueError("self_similarity needs x and y of the same shape")
        N, D = x.shape
        out = torch.empty(1, device=x.device, dtype=torch.float32)
        grad = torch.empty_like(x) if want_grad else None
        self._ck(self.lib.strotss_self_similarity(self._h, _ptr(x), x.stride(0), _ptr(y), y.stride(0), N, D, _ptr(out),
                                                  _ptr(grad), D, _stream(x.device)), "strotss_self_similarity")
        return out, grad

    def convert_rgb_to_yuv(self, x):
        x = _check_features("x", x)
        if x.dim() != 2 or x.shape[1] < 3:
            raise ValueError("convert_rgb_to_yuv expects (n, >=3)")
        if x.stride(1) != 1:
            x = x.contiguous()
        out = torch.empty(x.shape[0], 3, device=x.device, dtype=torch.float32)
        self._ck(self.lib.strotss_convert_rgb_to_yuv(self._h, _ptr(x), x.stride(0), x.shape[0], _ptr(out), _stream(x.device)),
                 "strotss_convert_rgb_to_yuv")
        return out

    def debug_gemm(self, A, B, alpha=1.0, tile_n=256):
        A = _check_features("A", A).contiguous()
        B = _check_features("B", B).contiguous()
        m, k = A.shape
        n = B.shape[0]
        Cm = torch.empty(m, n, device=A.device, dtype=torch.float32)
        self._ck(self.lib.strotss_debug_gemm(self._h, _ptr(A), m, _ptr(B), n, k, float(alpha), _ptr(Cm), tile_n, _stream(A.device)),
                 "strotss_debug_gemm")
        return Cm

    def debug_gemm_ta(self, At, B, alpha=1.0, C=None, variant=0):
        """C (+)= alpha * At^T @ B^T with At given (k, m): exercises the MN-major A descriptor path.  variant 0: single-CTA
        kernel; 1: CTA-pair kernel, B K-major; 2: CTA-pair kernel computing the Gram matrix At^T At (both operands MN-major;
        B is ignored except for its row count, which must equal m)."""
        At = _check_features("At", At).contiguous()
        B = _check_features("B", B).contiguous()
        k, m = At.shape
        n = B.shape[0]
        acc = C is not None
        if C is None:
            C = torch.empty(m, n, device=At.device, dtype=torch.float32)
        flags = (1 if acc else 0) | (2 if variant == 1 else 0) | (4 if variant == 2 else 0)
        self._ck(self.lib.strotss_debug_gemm_ta(self._h, _ptr(At), m, _ptr(B), n, k, float(alpha), _ptr(C), flags,
                                                _stream(At.device)), "strotss_debug_gemm_ta")
        return C


_shared: Dict[int, Handle] = {}
_pool: Dict[int, list] = {}


def acquire_handle(device: torch.device) -> Handle:
    """A handle for a loss module.  The reference builds a new StyleLoss per scale (run_strotss.py:100,128); handles
    released by collected modules are reused so that their device workspace is not re-allocated every scale."""
    if device.type != "cuda":
        raise RuntimeError(f"tensor is on {device}: strotss_tensorflow_b200 runs on CUDA sm_100 only (no CPU fallback)")
    idx = device.index if device.index is not None else torch.cuda.current_device()
    free = _pool.get(idx)
    if free:
        return free.pop()
    return Handle(torch.device("cuda", idx))


def release_handle(h: Optional[Handle]) -> None:
    if h is None or getattr(h, "_h", None) is None or _pool is None:
        return
    if h.world != 1:            # attached to a communicator: not reusable by an unrelated module
        return
    _pool.setdefault(h.index, []).append(h)


def shared_handle(device: torch.device) -> Handle:
    """Per-device handle used by the stateless functions of losses.py."""
    if device.type != "cuda":
        raise RuntimeError(f"tensor is on {device}: strotss_tensorflow_b200 runs on CUDA sm_100 only (no CPU fallback)")
    idx = device.index if device.index is not None else torch.cuda.current_device()
    h = _shared.get(idx)
    if h is None:
        h = Handle(torch.device("cuda", idx))
        _shared[idx] = h
    return h
