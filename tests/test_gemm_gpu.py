"""GPU parity: tcgen05 split-bf16 GEMM (through the C ABI) vs torch fp64 matmul."""
import ctypes

import pytest
import torch

from markushgrapher_b200 import _lib

pytestmark = pytest.mark.gpu


def run_gemm(M, N, K, bias=False, residual=False, act=0, planes=2, block_n=128, ksplit=1, swap=False, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    a = torch.randn(M, K, generator=g).cuda()
    b = torch.randn(N, K, generator=g).cuda()
    bi = torch.randn(N, generator=g).cuda() if bias else None
    res = torch.randn(M, N, generator=g).cuda() if residual else None
    c = torch.empty((N, M) if swap else (M, N), device="cuda")
    if swap and res is not None:
        res_in = res.t().contiguous()
    else:
        res_in = res
    rc = _lib.lib().mg_op_gemm(_lib.cur_stream(), M, N, K, _lib.ptr(a), _lib.ptr(b), _lib.ptr(c), _lib.ptr(bi),
                               _lib.ptr(res_in), act, planes, block_n, ksplit, int(swap))
    _lib.check(rc, "mg_op_gemm")
    torch.cuda.synchronize()
    ref = a.double() @ b.double().t()
    if bias:
        ref = ref + bi.double()
    if act == 1:
        ref = torch.relu(ref)
    elif act == 2:
        ref = torch.nn.functional.gelu(ref)
    if residual:
        ref = ref + res.double()
    out = c.t() if swap else c
    err = (out.double() - ref).norm() / ref.norm()
    maxerr = (out.double() - ref).abs().max().item()
    return err.item(), maxerr


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (256, 384, 128), (300, 200, 1024), (144, 96, 48), (1000, 33201 // 8, 1024)])
def test_gemm_split_fp32_accuracy(M, N, K):
    err, maxerr = run_gemm(M, N, K)
    assert err < 2e-5, (err, maxerr)


def test_gemm_bf16_plane():
    err, _ = run_gemm(256, 256, 512, planes=1)
    assert err < 1e-2
    assert err > 1e-5  # really bf16


@pytest.mark.parametrize("block_n", [32, 64, 128])
def test_gemm_block_n(block_n):
    err, _ = run_gemm(257, 100, 320, block_n=block_n)
    assert err < 2e-5


def test_gemm_epilogues():
    assert run_gemm(200, 256, 256, bias=True, act=1)[0] < 2e-5
    assert run_gemm(200, 256, 256, bias=True, act=2)[0] < 2e-5
    assert run_gemm(200, 256, 256, residual=True)[0] < 2e-5
    assert run_gemm(130, 70, 256, bias=True, residual=True)[0] < 2e-5


def test_gemm_splitk_swapped():
    # decode shape: M = output features, N = batch rows, transposed output, split-K with atomics
    assert run_gemm(1024, 32, 1024, block_n=32, ksplit=8, swap=True, residual=True)[0] < 2e-5
    assert run_gemm(4096, 32, 1024, block_n=32, ksplit=4, swap=True)[0] < 2e-5
    assert run_gemm(1024, 48, 4096, block_n=64, ksplit=16, swap=True)[0] < 2e-5
