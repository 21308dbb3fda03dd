"""The XLA FFI shim (ffi/b200sim_ffi.cc) as real, compiled source: it builds against the stand-in header when jaxlib is absent,
exports one handler per target, reports bad calls as XLA_FFI_Error, and -- driven here through hand-built call frames, the way
XLA's runtime would -- its GAE handler returns what ``b200sim_gae`` returns."""

import ctypes
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
HANDLERS = ("b200sim_ffi_engine_reset", "b200sim_ffi_engine_step", "b200sim_ffi_step_ctrl", "b200sim_ffi_rewards", "b200sim_ffi_gae")
c_sz, c_vp, c_i64 = ctypes.c_size_t, ctypes.c_void_p, ctypes.c_int64


class Buffer(ctypes.Structure):
    _fields_ = [("struct_size", c_sz), ("ext", c_vp), ("dtype", ctypes.c_int), ("data", c_vp), ("rank", c_i64), ("dims", ctypes.POINTER(c_i64))]


class Args(ctypes.Structure):
    _fields_ = [("struct_size", c_sz), ("ext", c_vp), ("size", c_i64), ("types", ctypes.POINTER(ctypes.c_int)), ("args", ctypes.POINTER(c_vp))]


class ByteSpan(ctypes.Structure):
    _fields_ = [("ptr", ctypes.c_char_p), ("len", c_sz)]


class Scalar(ctypes.Structure):
    _fields_ = [("dtype", ctypes.c_int), ("value", c_vp)]


class Attrs(ctypes.Structure):
    _fields_ = [("struct_size", c_sz), ("ext", c_vp), ("size", c_i64), ("types", ctypes.POINTER(ctypes.c_int)),
                ("names", ctypes.POINTER(ctypes.POINTER(ByteSpan))), ("attrs", ctypes.POINTER(c_vp))]


class ApiVersion(ctypes.Structure):
    _fields_ = [("struct_size", c_sz), ("ext", c_vp), ("major", ctypes.c_int), ("minor", ctypes.c_int)]


class ErrorCreateArgs(ctypes.Structure):
    _fields_ = [("struct_size", c_sz), ("ext", c_vp), ("message", ctypes.c_char_p), ("errc", ctypes.c_int)]


class StreamGetArgs(ctypes.Structure):
    _fields_ = [("struct_size", c_sz), ("ext", c_vp), ("ctx", c_vp), ("stream", c_vp)]


ERROR_CREATE = ctypes.CFUNCTYPE(c_vp, ctypes.POINTER(ErrorCreateArgs))
STREAM_GET = ctypes.CFUNCTYPE(c_vp, ctypes.POINTER(StreamGetArgs))


class Api(ctypes.Structure):
    _fields_ = [("struct_size", c_sz), ("ext", c_vp), ("version", ApiVersion), ("internal", c_vp), ("error_create", ERROR_CREATE),
                ("error_get_message", c_vp), ("error_destroy", c_vp), ("handler_register", c_vp), ("stream_get", STREAM_GET)]


class CallFrame(ctypes.Structure):
    _fields_ = [("struct_size", c_sz), ("ext", c_vp), ("api", ctypes.POINTER(Api)), ("ctx", c_vp), ("stage", ctypes.c_int), ("args", Args),
                ("rets", Args), ("attrs", Attrs), ("future", c_vp)]


F32, U8, S32, EXECUTE, BUFFER, SCALAR = 11, 6, 4, 3, 1, 3


class Frame:
    """One call frame + everything it points to, kept alive together."""

    def __init__(self, args, rets, attrs, stream: int = 0) -> None:  # noqa: ANN001
        self.errors: list[tuple[int, bytes]] = []
        self.keep: list = []

        def on_error(a):  # noqa: ANN001, ANN202
            self.errors.append((a.contents.errc, a.contents.message))
            return 0xDEAD

        def on_stream(a):  # noqa: ANN001, ANN202
            a.contents.stream = stream
            return None

        self.cb = (ERROR_CREATE(on_error), STREAM_GET(on_stream))
        self.api = Api(ctypes.sizeof(Api), None, ApiVersion(ctypes.sizeof(ApiVersion), None, 0, 1), None, self.cb[0], None, None, None, self.cb[1])
        self.frame = CallFrame(ctypes.sizeof(CallFrame), None, ctypes.pointer(self.api), None, EXECUTE, self._bufs(args), self._bufs(rets),
                               self._attrs(attrs), None)

    def _bufs(self, bufs) -> Args:  # noqa: ANN001
        n = len(bufs)
        types = (ctypes.c_int * max(n, 1))(*([BUFFER] * n))
        ptrs = (c_vp * max(n, 1))()
        for i, (dtype, ptr, shape) in enumerate(bufs):
            dims = (c_i64 * max(len(shape), 1))(*shape)
            b = Buffer(ctypes.sizeof(Buffer), None, dtype, ptr, len(shape), dims)
            self.keep += [dims, b]
            ptrs[i] = ctypes.addressof(b)
        self.keep += [types, ptrs]
        return Args(ctypes.sizeof(Args), None, n, types, ptrs)

    def _attrs(self, attrs) -> Attrs:  # noqa: ANN001
        n = len(attrs)
        types = (ctypes.c_int * max(n, 1))(*([SCALAR] * n))
        names = (ctypes.POINTER(ByteSpan) * max(n, 1))()
        vals = (c_vp * max(n, 1))()
        for i, (name, dtype, value) in enumerate(attrs):
            span = ByteSpan(name.encode(), len(name))
            sc = Scalar(dtype, ctypes.addressof(value))
            self.keep += [span, sc, value]
            names[i] = ctypes.pointer(span)
            vals[i] = ctypes.addressof(sc)
        self.keep += [types, names, vals]
        return Attrs(ctypes.sizeof(Attrs), None, n, types, names, vals)


@pytest.fixture(scope="module")
def shim() -> ctypes.CDLL:
    import __graft_entry__ as g

    g.build_cuda()
    subprocess.run(["make", "-C", str(ROOT / "ffi")], check=True, capture_output=True)
    lib = ctypes.CDLL(str(ROOT / "ffi" / "libb200sim_ffi.so"))
    for h in HANDLERS:
        fn = getattr(lib, h)
        fn.argtypes, fn.restype = [ctypes.POINTER(CallFrame)], c_vp
    return lib


def test_shim_builds_and_exports_every_handler(shim: ctypes.CDLL) -> None:
    assert shim.b200sim_ffi_built_against_stub() in (0, 1)
    out = subprocess.run(["nm", "-D", "--defined-only", str(ROOT / "ffi" / "libb200sim_ffi.so")], capture_output=True, text=True, check=True).stdout
    for h in HANDLERS:
        assert f" T {h}" in out


def test_bad_calls_come_back_as_xla_errors(shim: ctypes.CDLL) -> None:
    fr = Frame([], [], [])                       # no operands, no attributes
    assert shim.b200sim_ffi_gae(ctypes.byref(fr.frame)) == 0xDEAD
    assert fr.errors and fr.errors[0][0] == 3 and b"missing operand values" in fr.errors[0][1]
    fr = Frame([], [], [])
    assert shim.b200sim_ffi_engine_step(ctypes.byref(fr.frame)) == 0xDEAD and b"missing attribute model" in fr.errors[0][1]
    fr = Frame([], [], [])
    fr.frame.stage = 1                           # PREPARE: nothing to do, no error
    assert shim.b200sim_ffi_engine_step(ctypes.byref(fr.frame)) is None and not fr.errors
    bad = Frame([(S32, 0, (2, 3))] * 4, [(F32, 0, (2, 3))] * 3, [])   # wrong element type
    assert shim.b200sim_ffi_gae(ctypes.byref(bad.frame)) == 0xDEAD and b"wrong element type" in bad.errors[0][1]


@pytest.mark.gpu
def test_gae_handler_equals_the_c_abi_call(shim: ctypes.CDLL, walking_spec, walking_randomizers) -> None:  # noqa: ANN001
    import torch

    from ksim_b200.engine import B200Engine

    eng = B200Engine(walking_spec, 4, device="cuda:0", randomizers=walking_randomizers)
    b, t = 37, 24
    gen = torch.Generator(device="cuda:0").manual_seed(5)
    values, rewards = torch.randn((b, t), generator=gen, device="cuda:0"), torch.randn((b, t), generator=gen, device="cuda:0")
    dones = (torch.rand((b, t), generator=gen, device="cuda:0") < 0.1).to(torch.uint8)
    succ = (dones * (torch.rand((b, t), generator=gen, device="cuda:0") < 0.5)).to(torch.uint8)
    want = eng.gae(values, rewards, dones, succ, 0.97, 0.95)
    outs = [torch.zeros((b, t), device="cuda:0") for _ in range(3)]
    stream = torch.cuda.current_stream().cuda_stream
    fr = Frame([(F32, values.data_ptr(), (b, t)), (F32, rewards.data_ptr(), (b, t)), (U8, dones.data_ptr(), (b, t)), (U8, succ.data_ptr(), (b, t))],
               [(F32, o.data_ptr(), (b, t)) for o in outs],
               [("gamma", F32, ctypes.c_float(0.97)), ("lam", F32, ctypes.c_float(0.95)), ("flags", S32, ctypes.c_int32(0))], stream=stream)
    assert shim.b200sim_ffi_gae(ctypes.byref(fr.frame)) is None, fr.errors
    torch.cuda.synchronize()
    for got, ref in zip(outs, want):
        torch.testing.assert_close(got, ref, rtol=0, atol=0)


@pytest.mark.gpu
def test_engine_step_handler_equals_the_c_abi_call(shim: ctypes.CDLL, walking_spec, walking_randomizers) -> None:  # noqa: ANN001
    """Operand != result buffers (no aliasing): the handler copies the state over and steps the copy."""
    import torch

    from ksim_b200.engine import B200Engine

    n = 40
    eng = B200Engine(walking_spec, n, device="cuda:0", randomizers=walking_randomizers, seed=2)
    action = 0.3 * torch.randn((n, eng.NA), device="cuda:0")
    rng = torch.randint(-2**31, 2**31 - 1, (n, 2), dtype=torch.int32, device="cuda:0")
    state_in, keys_in = eng.state.clone(), eng.keys.clone()
    want = eng.engine_step(action, rng)
    state_o, keys_o = torch.zeros_like(state_in), torch.zeros_like(keys_in)
    data_o = torch.zeros_like(want.block)
    handle = ctypes.c_uint64(eng._handle.value)
    U32, U64 = 8, 9
    fr = Frame([(F32, action.data_ptr(), tuple(action.shape)), (U32, rng.data_ptr(), (n, 2)), (F32, eng.rand.data_ptr(), tuple(eng.rand.shape)),
                (F32, state_in.data_ptr(), tuple(state_in.shape)), (U32, keys_in.data_ptr(), tuple(keys_in.shape))],
               [(F32, state_o.data_ptr(), tuple(state_o.shape)), (U32, keys_o.data_ptr(), tuple(keys_o.shape)), (F32, data_o.data_ptr(), tuple(data_o.shape))],
               [("model", U64, handle)], stream=torch.cuda.current_stream().cuda_stream)
    assert shim.b200sim_ffi_engine_step(ctypes.byref(fr.frame)) is None, fr.errors
    torch.cuda.synchronize()
    torch.testing.assert_close(data_o, want.block, rtol=0, atol=0)
    torch.testing.assert_close(state_o, eng.state, rtol=0, atol=0)
