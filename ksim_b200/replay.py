"""Recurrent policy replay on the B200 tensor cores: the host side of ``b200sim_gru_forward`` / ``b200sim_gru_backward``.

Reference: the PPO update replays the policy over every minibatch trajectory (``ksim/task/ppo.py:417-488`` ->
``get_ppo_variables``, ``examples/walking.py:674-728``: an ``xax.scan`` of ``Actor.forward`` / ``Critic.forward``, each
``input_proj -> depth x eqx.nn.GRUCell(512) -> MLP``, with the carry reset where ``transition.done``).  Only
``h_{t-1} W_hh^T`` and the gate arithmetic are sequential; ``x_t W_ih^T`` for all ``t`` is one batched GEMM.  ``GRUStack``
keeps the reference's parameterisation (``eqx.nn.GRUCell``: ``weight_ih [3H, I]``, ``weight_hh [3H, H]``, ``bias [3H]``,
``bias_n [H]``; gate order r, z, n) and runs the sequential part of up to two stacks (actor and critic) side by side in
ONE persistent tcgen05 kernel per layer and direction; the batched GEMMs around it are plain library GEMMs (bf16 operands,
fp32 accumulation) issued through torch.  ``gru_stack_reference`` is the fp32 restatement the tests compare against.

All sequence tensors are time-major ``[T, B, ...]`` at this module's interface (feature-major ``[features, T B]`` between
the kernels and the GEMMs inside).  There is no fallback: without ``libb200sim.so`` or off a CUDA device
the tensor-core path raises.
"""

from __future__ import annotations

__all__ = ["GRUStack", "gru_stacks_forward", "gru_stack_reference", "gru_cell_reference", "HIDDEN", "shadow"]

import ctypes
import math
from typing import Sequence

import torch
import torch.nn as nn

from . import binding
from . import codes as C

HIDDEN = C.B2S_GRU_HIDDEN


class _FwdProblem(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in ("w_hh", "gi", "b_i", "b_hn", "h0", "y", "hst", "h_last", "saved")]


class _BwdProblem(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in ("w_hh_t", "dy", "saved", "dgi", "dghn", "dh0")]


def _p(t: torch.Tensor | None) -> int | None:
    return None if t is None else t.data_ptr()


def _lib() -> ctypes.CDLL:
    return binding.lib()


def _check(rc: int, what: str) -> None:
    if rc != 0:
        msg = _lib().b200sim_gru_last_error()
        raise binding.B200SimError(f"{what} failed with code {rc}: {msg.decode() if msg else ''}")


_WORKSPACES: dict[tuple, torch.Tensor] = {}


def _workspace(n: int, t: int, b: int, backward: bool, device: torch.device) -> torch.Tensor:
    """Exchange buffers + flags of one launch, cached per shape so that captured CUDA graphs see fixed addresses."""
    key = (n, t, b, backward, device)
    ws = _WORKSPACES.get(key)
    if ws is None:
        ws = torch.empty(int(_lib().b200sim_gru_workspace_bytes(n, t, b, int(backward))), dtype=torch.uint8, device=device)
        _WORKSPACES[key] = ws
    return ws


def last_status(n: int, t: int, b: int, backward: bool, device: torch.device) -> int:
    """Device-side status word of the most recent launch of that shape (0 = ok, else ``B2S_GRU_ERR_*``).  Synchronises."""
    ws = _WORKSPACES.get((n, t, b, backward, torch.device(device)))
    return 0 if ws is None else int(ws[:4].view(torch.int32).item())


def shadow(p: torch.Tensor) -> torch.Tensor:
    """bf16 copy of a parameter for the tensor-core GEMMs: the shadow ``ksim_b200.optim.FlatParams`` keeps up to date inside
    the optimiser step (``p.shadow``), else a fresh cast."""
    s = getattr(p, "shadow", None)
    return s if s is not None else p.detach().to(torch.bfloat16)


_ONES: dict[tuple, torch.Tensor] = {}


def _ones_col(n: int, device: torch.device) -> torch.Tensor:
    key = (n, str(device))
    if key not in _ONES:
        _ONES[key] = torch.ones((n, 8), dtype=torch.bfloat16, device=device)   # 8 columns: a 16-byte-aligned operand
    return _ONES[key]


def _row_sums(x: torch.Tensor) -> torch.Tensor:
    """fp32 sums over the columns of a contiguous bf16 ``[rows, cols]`` matrix (``b200sim_row_sums``)."""
    rows, cols = int(x.shape[0]), int(x.shape[1])
    out = torch.empty((rows,), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _check(_lib().b200sim_row_sums(ctypes.c_void_p(torch.cuda.current_stream(x.device).cuda_stream), rows, cols, _p(x), _p(out)),
               "b200sim_row_sums")
    return out


def _mm(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """bf16 x bf16 -> fp32 library GEMM."""
    return torch.mm(a, b, out_dtype=torch.float32)


class GRUStack(nn.Module):
    """``depth`` stacked ``eqx.nn.GRUCell(hidden, hidden)`` (examples/walking.py:78-91), same parameter shapes and init
    (uniform in ``+-1/sqrt(hidden)``)."""

    def __init__(self, depth: int, hidden: int = HIDDEN, input_size: int | None = None) -> None:
        super().__init__()
        if hidden != HIDDEN:
            raise ValueError(f"the tensor-core kernels are built for hidden = {HIDDEN}")
        self.depth, self.hidden = depth, hidden
        lim = 1.0 / math.sqrt(hidden)
        mk = lambda *s: nn.Parameter(torch.empty(*s).uniform_(-lim, lim))  # noqa: E731
        self.weight_ih = nn.ParameterList([mk(3 * hidden, (input_size or hidden) if i == 0 else hidden) for i in range(depth)])
        self.weight_hh = nn.ParameterList([mk(3 * hidden, hidden) for _ in range(depth)])
        self.bias = nn.ParameterList([mk(3 * hidden) for _ in range(depth)])
        self.bias_n = nn.ParameterList([mk(hidden) for _ in range(depth)])

    def layer_params(self) -> list[torch.Tensor]:
        out: list[torch.Tensor] = []
        for i in range(self.depth):
            out += [self.weight_ih[i], self.weight_hh[i], self.bias[i], self.bias_n[i]]
        return out

    def forward(self, x: torch.Tensor, carry: torch.Tensor | None, keep: torch.Tensor | None) -> tuple[torch.Tensor, torch.Tensor]:
        (y, h_last), = gru_stacks_forward([self], [x], [carry], keep)
        return y, h_last


def gru_cell_reference(x: torch.Tensor, h: torch.Tensor, w_ih: torch.Tensor, w_hh: torch.Tensor, bias: torch.Tensor,
                       bias_n: torch.Tensor) -> torch.Tensor:
    """``eqx.nn.GRUCell.__call__`` in plain torch (any dtype): the restatement the kernels are tested against."""
    gi = x @ w_ih.T + bias
    gh = h @ w_hh.T
    hd = h.shape[-1]
    r = torch.sigmoid(gi[..., :hd] + gh[..., :hd])
    z = torch.sigmoid(gi[..., hd : 2 * hd] + gh[..., hd : 2 * hd])
    n = torch.tanh(gi[..., 2 * hd :] + r * (gh[..., 2 * hd :] + bias_n))
    return n + z * (h - n)


def gru_stack_reference(stack: GRUStack, x: torch.Tensor, carry: torch.Tensor | None, keep: torch.Tensor | None,
                        ) -> tuple[torch.Tensor, torch.Tensor]:
    """Step-by-step scan of the stack as the reference runs it (examples/walking.py:121-127, 715-719): returns the top layer's
    outputs ``[T, B, H]`` and the carry after the last step ``[depth, B, H]``, in the dtype of ``x``."""
    t_, b_ = x.shape[0], x.shape[1]
    h = [torch.zeros(b_, stack.hidden, dtype=x.dtype, device=x.device) if carry is None else carry[i].to(x.dtype) for i in range(stack.depth)]
    ys = []
    for t in range(t_):
        v = x[t]
        for i in range(stack.depth):
            v = gru_cell_reference(v, h[i], stack.weight_ih[i].to(x.dtype), stack.weight_hh[i].to(x.dtype), stack.bias[i].to(x.dtype),
                                   stack.bias_n[i].to(x.dtype))
            h[i] = v if keep is None else v * keep[t].to(x.dtype)[:, None]
        ys.append(v)
    return torch.stack(ys), torch.stack(h)


# Measurement hook (bench.py): a list here makes every eager GRU kernel launch append (direction, T, B, stacks, start event, end
# event), CUDA events recorded on the launch stream around the one launch.  None (the default) costs nothing.
KERNEL_TIMES: list | None = None


def _timed_launch(direction: str, t_: int, b_: int, n_stacks: int, dev: torch.device, launch) -> None:  # noqa: ANN001
    if KERNEL_TIMES is None or torch.cuda.is_current_stream_capturing():
        launch()
        return
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(torch.cuda.current_stream(dev))
    launch()
    e1.record(torch.cuda.current_stream(dev))
    KERNEL_TIMES.append((direction, t_, b_, n_stacks, e0, e1))


class _GruStacksFn(torch.autograd.Function):
    """Up to ``B2S_GRU_MAX_PROBLEMS`` GRU stacks of equal depth over the same ``(T, B)`` batch, layer by layer; each layer of
    all stacks is one persistent kernel launch per direction.  Between the kernels and the library GEMMs every sequence array
    is FEATURE-major (``[features, T B]``): coalesced for the kernels (one thread per batch row), a transposed operand for the
    GEMMs, and row sums for the bias gradients."""

    @staticmethod
    def forward(ctx, n_stacks: int, depth: int, keep: torch.Tensor | None, want_grad: bool, *args: torch.Tensor):  # noqa: ANN001, ANN205
        lib = _lib()
        xs, h0s = args[:n_stacks], args[n_stacks : 2 * n_stacks]
        params = args[2 * n_stacks :]  # per stack, per layer: w_ih, w_hh, bias, bias_n
        t_, b_ = int(xs[0].shape[0]), int(xs[0].shape[1])
        tb = t_ * b_
        dev = xs[0].device
        if dev.type != "cuda":
            raise binding.B200SimError("the GRU replay kernels need CUDA tensors; there is no CPU path")
        hd = HIDDEN
        if keep is not None:
            keep = keep.to(torch.float32).contiguous()
        stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        n_saved = int(lib.b200sim_gru_saved_floats(t_, b_))
        # layer inputs as [I, T B] operands: the first is a transposed view of the row-major input, the others are the
        # feature-major outputs of the layer below
        cur = [x.detach().to(torch.bfloat16).reshape(tb, -1).t() for x in xs]
        saved_layers = []
        h_last = [torch.empty((depth, b_, hd), dtype=torch.float32, device=dev) for _ in range(n_stacks)]
        ws = _workspace(n_stacks, t_, b_, False, dev)
        for layer in range(depth):
            probs = (_FwdProblem * n_stacks)()
            hold = []
            nxt = []
            for s in range(n_stacks):
                w_ih, w_hh, bias, bias_n = params[(s * depth + layer) * 4 : (s * depth + layer) * 4 + 4]
                w_ih16, w_hh16 = shadow(w_ih), shadow(w_hh).contiguous()
                gi = _mm(w_ih16, cur[s])  # [3H, T B] fp32
                y = torch.empty((hd, tb), dtype=torch.bfloat16, device=dev)
                hst = torch.empty((hd, tb), dtype=torch.bfloat16, device=dev)
                sv = torch.empty(n_saved, dtype=torch.float32, device=dev) if want_grad else None
                h0 = None if h0s[s] is None else h0s[s][layer].to(torch.float32).contiguous()
                bi, bn = bias.detach().float().contiguous(), bias_n.detach().float().contiguous()
                hl = h_last[s][layer]
                p = probs[s]
                p.w_hh, p.gi, p.b_i, p.b_hn, p.h0, p.y, p.hst, p.h_last, p.saved = (_p(w_hh16), _p(gi), _p(bi), _p(bn), _p(h0), _p(y), _p(hst),
                                                                                     _p(hl), _p(sv))
                hold += [w_hh16, gi, bi, bn, h0]
                nxt.append(y)
                saved_layers.append((cur[s], w_ih16, w_hh16, hst, sv))
            with torch.cuda.device(dev):
                _timed_launch("fwd", t_, b_, n_stacks, dev, lambda: _check(
                    lib.b200sim_gru_forward(stream, n_stacks, ctypes.byref(probs), t_, b_, _p(keep), _p(ws), ws.numel()), "b200sim_gru_forward"))
            del hold  # stream-ordered allocator: the memory is not reused before the launch has run
            cur = nxt
        ctx.n_stacks, ctx.depth, ctx.shape, ctx.keep = n_stacks, depth, (t_, b_), keep
        ctx.saved_layers = saved_layers
        ctx.x_dtypes = [x.dtype for x in xs]
        ctx.x_shapes = [tuple(x.shape) for x in xs]
        outs = []
        for s in range(n_stacks):
            outs += [cur[s].t().unflatten(0, (t_, b_)), h_last[s]]  # [T, B, H] view of the feature-major top output
            ctx.mark_non_differentiable(h_last[s])
        return tuple(outs)

    @staticmethod
    def backward(ctx, *grads: torch.Tensor):  # noqa: ANN205
        lib = _lib()
        n_stacks, depth, (t_, b_), keep = ctx.n_stacks, ctx.depth, ctx.shape, ctx.keep
        tb = t_ * b_
        hd = HIDDEN
        dev = ctx.saved_layers[0][0].device
        stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        ws = _workspace(n_stacks, t_, b_, True, dev)
        dy = []  # feature-major fp32 [H, T B]
        for s in range(n_stacks):
            g = grads[2 * s]
            dy.append(torch.zeros((hd, tb), dtype=torch.float32, device=dev) if g is None
                      else g.reshape(tb, hd).t().to(torch.float32).contiguous())
        pgrads: list[list[torch.Tensor | None]] = [[None] * (4 * depth) for _ in range(n_stacks)]
        dx: list[torch.Tensor | None] = [None] * n_stacks
        for layer in reversed(range(depth)):
            probs = (_BwdProblem * n_stacks)()
            outs = []
            hold = []
            for s in range(n_stacks):
                x_in, w_ih16, w_hh16, hst, sv = ctx.saved_layers[layer * n_stacks + s]
                w_hh_t = w_hh16.t().contiguous()
                dgi = torch.empty((3 * hd, tb), dtype=torch.bfloat16, device=dev)
                dghn = torch.empty((hd, tb), dtype=torch.bfloat16, device=dev)
                p = probs[s]
                p.w_hh_t, p.dy, p.saved, p.dgi, p.dghn, p.dh0 = _p(w_hh_t), _p(dy[s]), _p(sv), _p(dgi), _p(dghn), None
                outs.append((dgi, dghn))
                hold += [w_hh_t, dy[s]]
            with torch.cuda.device(dev):
                _timed_launch("bwd", t_, b_, n_stacks, dev, lambda: _check(
                    lib.b200sim_gru_backward(stream, n_stacks, ctypes.byref(probs), t_, b_, _p(keep), _p(ws), ws.numel()), "b200sim_gru_backward"))
            del hold
            for s in range(n_stacks):
                x_in, w_ih16, w_hh16, hst, sv = ctx.saved_layers[layer * n_stacks + s]
                dgi, dghn = outs[s]
                hst_r = hst.t()  # [T B, H] operand
                g_w_ih = _mm(dgi, x_in.t())
                g_w_hh = torch.cat([_mm(dgi[: 2 * hd], hst_r), _mm(dghn, hst_r)], dim=0)
                g_b, g_bn = _row_sums(dgi), _row_sums(dghn)   # bias gradients: HBM-bound row sums of the gate gradients
                pgrads[s][layer * 4 : layer * 4 + 4] = [g_w_ih, g_w_hh, g_b, g_bn]
                if layer > 0:
                    dy[s] = _mm(w_ih16.t(), dgi)  # [H, T B]: gradient w.r.t. the layer below's feature-major output
                else:
                    dx[s] = _mm(dgi.t(), w_ih16).view(ctx.x_shapes[s]).to(ctx.x_dtypes[s])  # row-major, like the input
        flat: list[torch.Tensor | None] = []
        for s in range(n_stacks):
            flat += pgrads[s]
        return (None, None, None, None, *dx, *([None] * n_stacks), *flat)


def gru_stacks_forward(stacks: Sequence[GRUStack], xs: Sequence[torch.Tensor], carries: Sequence[torch.Tensor | None],
                       keep: torch.Tensor | None) -> list[tuple[torch.Tensor, torch.Tensor]]:
    """Runs ``stacks[i]`` over ``xs[i] [T, B, I_i]`` from ``carries[i] [depth, B, H]`` (``None`` = zeros), all stacks side by side.
    ``keep [T, B]``: 1 where the episode continues after step t, 0 where it ended (the carry restarts from zero).
    Returns per stack ``(y [T, B, H] bf16, carry_after [depth, B, H] fp32)``; differentiable in ``xs`` and the parameters."""
    n = len(stacks)
    if not 1 <= n <= C.B2S_GRU_MAX_PROBLEMS or len(xs) != n or len(carries) != n:
        raise ValueError(f"1..{C.B2S_GRU_MAX_PROBLEMS} stacks with one input and one carry each")
    depth = stacks[0].depth
    if any(s.depth != depth for s in stacks):
        raise ValueError("stacks run side by side must have the same depth")
    params: list[torch.Tensor] = []
    for s in stacks:
        params += s.layer_params()
    want_grad = torch.is_grad_enabled() and (any(p.requires_grad for p in params) or any(x.requires_grad for x in xs))
    res = _GruStacksFn.apply(n, depth, keep, want_grad, *xs, *carries, *params)
    return [(res[2 * i], res[2 * i + 1]) for i in range(n)]
