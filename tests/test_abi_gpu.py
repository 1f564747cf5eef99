"""The C ABI itself, called with raw pointers (ctypes), including its error codes."""
import ctypes

import pytest
import torch

from tests.util import assert_close

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def L():
    import diffcodec_b200
    return diffcodec_b200._lib


def test_raw_call_and_error_codes(L):
    lib = L.lib()
    tin = torch.rand(1, 3, 16, 16, device="cuda"); flow = torch.zeros(1, 2, 16, 16, device="cuda")
    out = torch.empty_like(tin)
    st = torch.cuda.current_stream().cuda_stream
    need = lib.dcb_splat_workspace_bytes(1, 3, 16, 16, L.DCB_F32, L.MODE_AVG, 0)
    assert need >= 16 * 16 * 16
    ws = torch.empty(need, dtype=torch.uint8, device="cuda")
    before = lib.dcb_launch_count()
    rc = lib.dcb_splat_fwd(L.desc(tin), L.desc(flow), None, L.desc(out), None, None, ws.data_ptr(), need, L.MODE_AVG, L.EPS_ADD, 0, st)
    assert rc == 0 and lib.dcb_launch_count() - before == 1          # a small frame: ONE launch (cluster kernel; the memset is not a kernel of ours)
    torch.cuda.synchronize()
    assert_close(out, tin / (1 + 1e-7), 1e-6, "raw avg")
    # workspace too small / missing
    assert lib.dcb_splat_fwd(L.desc(tin), L.desc(flow), None, L.desc(out), None, None, ws.data_ptr(), 16, L.MODE_AVG, 0, 0, st) == L.E_WORKSPACE
    assert b"workspace" in lib.dcb_last_error()
    assert lib.dcb_splat_fwd(L.desc(tin), L.desc(flow), None, L.desc(out), None, None, None, 0, L.MODE_AVG, 0, 0, st) == L.E_WORKSPACE
    # required pointer missing
    assert lib.dcb_splat_fwd(None, L.desc(flow), None, L.desc(out), None, None, None, 0, L.MODE_SUM, 0, 0, st) == L.E_NULL
    # metric rules (softsplat.py:235-238)
    m = torch.zeros(1, 1, 16, 16, device="cuda")
    assert lib.dcb_splat_fwd(L.desc(tin), L.desc(flow), L.desc(m), L.desc(out), None, None, ws.data_ptr(), need, L.MODE_SUM, 0, 0, st) == L.E_MODE
    assert lib.dcb_splat_fwd(L.desc(tin), L.desc(flow), None, L.desc(out), None, None, ws.data_ptr(), need, L.MODE_SOFT, 0, 0, st) == L.E_MODE
    assert lib.dcb_splat_fwd(L.desc(tin), L.desc(flow), None, L.desc(out), None, None, ws.data_ptr(), need, 9, 0, 0, st) == L.E_MODE
    # shape / dtype
    bad = torch.zeros(1, 2, 16, 17, device="cuda")
    assert lib.dcb_splat_fwd(L.desc(tin), L.desc(bad), None, L.desc(out), None, None, ws.data_ptr(), need, L.MODE_SUM, 0, 0, st) == L.E_SHAPE
    assert lib.dcb_splat_fwd(L.desc(tin), L.desc(flow.double()), None, L.desc(out), None, None, ws.data_ptr(), need, L.MODE_SUM, 0, 0, st) == L.E_DTYPE
    assert lib.dcb_splat_fwd(L.desc(tin), L.desc(flow), None, L.desc(out.permute(0, 1, 3, 2)), None, None, ws.data_ptr(), need, L.MODE_SUM, 0, 0, st) == L.E_SHAPE
    torch.cuda.synchronize()


def test_ws_clean_protocol(L):
    """DCB_FLAG_WS_CLEAN: a zeroed workspace stays zeroed, and repeated calls agree."""
    lib = L.lib()
    tin = torch.rand(2, 3, 32, 48, device="cuda"); flow = torch.randn(2, 2, 32, 48, device="cuda") * 3
    m = torch.randn(2, 1, 32, 48, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    need = lib.dcb_splat_workspace_bytes(2, 3, 32, 48, L.DCB_F32, L.MODE_SOFT, 0)
    ws = torch.zeros(need, dtype=torch.uint8, device="cuda")
    outs = []
    for _ in range(3):
        out = torch.empty_like(tin)
        assert lib.dcb_splat_fwd(L.desc(tin), L.desc(flow), L.desc(m), L.desc(out), None, None, ws.data_ptr(), need, L.MODE_SOFT, 0, L.FLAG_WS_CLEAN, st) == 0
        outs.append(out)
    torch.cuda.synchronize()
    assert int(ws.count_nonzero()) == 0
    assert_close(outs[1], outs[0], 1e-6, "ws clean repeat"); assert_close(outs[2], outs[0], 1e-6, "ws clean repeat")
    # planar layout too (C+1 > 4)
    t8 = torch.rand(1, 8, 16, 16, device="cuda"); f8 = torch.randn(1, 2, 16, 16, device="cuda"); m8 = torch.zeros(1, 1, 16, 16, device="cuda")
    need = lib.dcb_splat_workspace_bytes(1, 8, 16, 16, L.DCB_F32, L.MODE_SOFT, 0)
    ws = torch.zeros(need, dtype=torch.uint8, device="cuda"); o8 = torch.empty_like(t8)
    assert lib.dcb_splat_fwd(L.desc(t8), L.desc(f8), L.desc(m8), L.desc(o8), None, None, ws.data_ptr(), need, L.MODE_SOFT, 0, L.FLAG_WS_CLEAN, st) == 0
    torch.cuda.synchronize()
    assert int(ws.count_nonzero()) == 0


def test_graph_capturable(L):
    """No allocation, sync or stream creation inside the library: a call can be captured in a CUDA graph."""
    lib = L.lib()
    tin = torch.rand(1, 3, 64, 64, device="cuda"); flow = torch.randn(1, 2, 64, 64, device="cuda")
    out = torch.empty_like(tin)
    need = lib.dcb_splat_workspace_bytes(1, 3, 64, 64, L.DCB_F32, L.MODE_AVG, 0)
    ws = torch.zeros(need, dtype=torch.uint8, device="cuda")
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(s):
        d = (L.desc(tin), L.desc(flow), L.desc(out))
        with torch.cuda.graph(g, stream=s):
            rc = lib.dcb_splat_fwd(d[0], d[1], None, d[2], None, None, ws.data_ptr(), need, L.MODE_AVG, 0, L.FLAG_WS_CLEAN, torch.cuda.current_stream().cuda_stream)
        assert rc == 0
    g.replay(); torch.cuda.synchronize()
    first = out.clone()
    tin.mul_(2.0)
    g.replay(); torch.cuda.synchronize()
    assert_close(out, first * 2, 1e-5, "graph replay")
