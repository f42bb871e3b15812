"""GPU tests (-m gpu) of the tcgen05 tensor-core path: the single-CTA UMMA self test (descriptor / layout semantics in
isolation) and the fused detector kernel (precision "bf16x3") against the fp32 kernel and the fp64 oracle."""
import ctypes

import numpy as np
import pytest
import torch

from oracle import net as onet
from tests.conftest import pkg
from tests.test_model_gpu import compare, run_pipeline

pytestmark = pytest.mark.gpu

# bf16x3 carries every operand of a tensor-core contraction as two bf16 terms (16 mantissa bits): errors are bounded by the
# "bf16x3" row of the one tolerance table, oracle/parity.py
P3 = "bf16x3"


def make_image(A, lbo, sbo):
    """canonical K-major no-swizzle bf16 operand image: element (r,k) at (k/8)*lbo + (r/8)*sbo + (r%8)*16 + (k%8)*2"""
    rows, K = A.shape
    r = torch.arange(rows).view(-1, 1)
    k = torch.arange(K).view(1, -1)
    off = (k // 8) * lbo + (r // 8) * sbo + (r % 8) * 16 + (k % 8) * 2
    nbytes = int(off.max()) + 2
    nbytes = (nbytes + 15) // 16 * 16
    img = torch.zeros(nbytes // 2, dtype=torch.int16)
    img[(off // 2).reshape(-1)] = A.to(torch.bfloat16).view(torch.int16).reshape(-1)
    return img, nbytes


def make_image_mn(B, lbo, sbo):
    """canonical MN-major no-swizzle bf16 operand image: element (n,k) at (n/8)*sbo + (k/8)*lbo + (k%8)*16 + (n%8)*2"""
    rows, K = B.shape
    n = torch.arange(rows).view(-1, 1)
    k = torch.arange(K).view(1, -1)
    off = (n // 8) * sbo + (k // 8) * lbo + (k % 8) * 16 + (n % 8) * 2
    nbytes = (int(off.max()) + 2 + 15) // 16 * 16
    img = torch.zeros(nbytes // 2, dtype=torch.int16)
    img[(off // 2).reshape(-1)] = B.to(torch.bfloat16).view(torch.int16).reshape(-1)
    return img, nbytes


@pytest.mark.parametrize("N,K,lbo_b,sbo_b", [(64, 64, 128, 1024), (64, 128, 128, 2048), (64, 16, 128, 256), (64, 64, 1024, 128)])
def test_umma_selftest_b_mn_major(cuda, N, K, lbo_b, sbo_b):
    """B operand stored sample-contiguous (MN-major): the layout the E1 epilogues write with 16-byte stores"""
    lib_mod = pkg("_lib")
    L = lib_mod.lib()
    g = torch.Generator().manual_seed(N * 77 + K)
    A = torch.randn((128, K), generator=g).to(torch.bfloat16).float()
    B = torch.randn((N, K), generator=g).to(torch.bfloat16).float()
    a_img, a_bytes = make_image(A, 2048, 128)
    b_img, b_bytes = make_image_mn(B, lbo_b, sbo_b)
    a_d, b_d = a_img.to(cuda), b_img.to(cuda)
    D = torch.full((128, N), float("nan"), device=cuda)
    rc = L.f3d_debug_umma_selftest(lib_mod.ptr(a_d), lib_mod.ptr(b_d), lib_mod.ptr(D), N, K, 2048, 128, lbo_b, sbo_b, a_bytes, b_bytes,
                                   1 | 2, lib_mod.stream())
    lib_mod.check(rc, "umma_selftest")
    torch.cuda.synchronize()
    want = A.double() @ B.double().t()
    err = (D.cpu().double() - want).abs().max().item()
    assert err < 1e-3 * max(1.0, want.abs().max().item()), "UMMA self test (B MN-major): max err %.3e" % err


@pytest.mark.parametrize("a_in_tmem", [0, 1])
@pytest.mark.parametrize("N,K,lbo_b", [(64, 64, 1024), (64, 128, 1040), (8, 16, 128), (256, 32, 4096), (64, 16, 1040)])
def test_umma_selftest(cuda, N, K, lbo_b, a_in_tmem):
    lib_mod = pkg("_lib")
    L = lib_mod.lib()
    g = torch.Generator().manual_seed(N * 1000 + K)
    A = torch.randn((128, K), generator=g).to(torch.bfloat16).float()
    B = torch.randn((N, K), generator=g).to(torch.bfloat16).float()
    a_img, a_bytes = make_image(A, 2048, 128)
    b_img, b_bytes = make_image(B, lbo_b, 128)
    a_d, b_d = a_img.to(cuda), b_img.to(cuda)
    D = torch.full((128, N), float("nan"), device=cuda)
    rc = L.f3d_debug_umma_selftest(lib_mod.ptr(a_d), lib_mod.ptr(b_d), lib_mod.ptr(D), N, K, 2048, 128, lbo_b, 128, a_bytes, b_bytes,
                                   a_in_tmem, lib_mod.stream())
    lib_mod.check(rc, "umma_selftest")
    torch.cuda.synchronize()
    want = A.double() @ B.double().t()
    err = (D.cpu().double() - want).abs().max().item()
    assert err < 1e-3 * max(1.0, want.abs().max().item()), "UMMA self test: max err %.3e" % err


def test_detector_bf16x3_vs_fp32_and_oracle(cuda):
    xyz = pkg("synth").make_batch(3, 4096, seed0=21)
    for rb in (True, False):
        params = onet.init_params(seed=2, randomize_bn=rb)
        out_tc, _ = run_pipeline(xyz, params, 200, precision="bf16x3")
        out_32, _ = run_pipeline(xyz, params, 200, precision="fp32")
        ref = onet.inference_model(xyz, onet.to_torch(params, torch.float64), num_clusters=200, dtype=torch.float64)
        assert torch.equal(out_tc["idx"], out_32["idx"])
        e = compare(out_tc, ref, P3, "bf16x3 rb=%s" % rb)
        print("bf16x3 errors (att rel, ori rad, feat abs):", e)


@pytest.mark.parametrize("F,no_regress", [(32, False), (64, False), (16, True), (128, False)])
def test_bf16x3_feature_dims(cuda, F, no_regress):
    """descriptor tensor kernel for feature_dim <= 64 (MID = 128); feature_dim 128 routes to the exact fp32 kernel"""
    xyz = pkg("synth").make_batch(2, 4096, seed0=31 + F)
    params = onet.init_params(seed=5, feature_dim=F, randomize_bn=True)
    out, _ = run_pipeline(xyz, params, 150, F=F, precision="bf16x3", no_regress=no_regress)
    ref = onet.inference_model(xyz, onet.to_torch(params, torch.float64), num_clusters=150, feature_dim=F,
                               no_regress=no_regress, dtype=torch.float64)
    e = compare(out, ref, P3, "bf16x3 F=%d" % F)
    print("bf16x3 F=%d errors:" % F, e)


def test_detector_bf16x3_c1(cuda):
    xyz = pkg("synth").base_cloud("oxford")[None]
    params = onet.init_params(seed=0, randomize_bn=True)
    out, _ = run_pipeline(xyz, params, 512, precision="bf16x3")
    ref = onet.inference_model(xyz, onet.to_torch(params, torch.float64), num_clusters=512, dtype=torch.float64)
    compare(out, ref, P3, "C1 bf16x3")
