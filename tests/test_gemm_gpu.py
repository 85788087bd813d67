"""GPU unit tests of the dense-layer primitives against plain PyTorch fp32 math.

bf16 kernels are compared with an fp32 matmul of the SAME bf16-rounded operands, so the
only differences are accumulation order and the final bf16 rounding of the output.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _bf16(x):
    return x.to(torch.bfloat16)


@pytest.mark.parametrize("M,N,K", [(128, 256, 256), (4096, 256, 256), (8192 + 64, 256, 64),
                                   (2048, 128, 256), (1024, 64, 256), (3 * 128 * 148 + 192, 256, 320)])
def test_gemm_bf16_plain(cuda_dev, M, N, K):
    from upnerf_b200 import _lib as L

    g = torch.Generator(device="cpu").manual_seed(M + N + K)
    A = _bf16(torch.randn(M, K, generator=g)).to(cuda_dev)
    B = _bf16(torch.randn(N, K, generator=g) / K ** 0.5).to(cuda_dev)
    Cout = torch.full((M, N), float("nan"), dtype=torch.bfloat16, device=cuda_dev)
    L.gemm_bf16(A, B, Cout, M, N, K)
    torch.cuda.synchronize()
    ref = A.float() @ B.float().t()
    err = (Cout.float() - ref).abs().max().item()
    assert err < 0.03, f"max abs err {err}"
    # bf16 rounding of an O(1) value is <= 2^-8 relative
    assert torch.allclose(Cout.float(), ref, rtol=1e-2, atol=1e-2)


def test_gemm_bf16_strided_views(cuda_dev):
    """A and C are column slices of wider buffers (the skip-concat buffer of layer 5)."""
    from upnerf_b200 import _lib as L

    M, N, K = 1024, 256, 64
    g = torch.Generator(device="cpu").manual_seed(3)
    Xbuf = _bf16(torch.randn(M, 320, generator=g)).to(cuda_dev)
    B = _bf16(torch.randn(N, K, generator=g) / 8).to(cuda_dev)
    Cbuf = torch.zeros(M, 320, dtype=torch.bfloat16, device=cuda_dev)
    A = Xbuf[:, 256:320]
    L.gemm_bf16(A, B, Cbuf[:, :256], M, N, K, lda=320, ldc=320)
    torch.cuda.synchronize()
    ref = A.float() @ B.float().t()
    assert torch.allclose(Cbuf[:, :256].float(), ref, rtol=1e-2, atol=1e-2)
    assert (Cbuf[:, 256:] == 0).all()


@pytest.mark.parametrize("aux_mode", [0, 1, 2])
def test_gemm_bf16_epilogue(cuda_dev, aux_mode):
    from upnerf_b200 import _lib as L

    M, N, K, S = 64 * 37, 256, 128, 64
    R = M // S
    g = torch.Generator(device="cpu").manual_seed(11 + aux_mode)
    A = _bf16(torch.randn(M, K, generator=g)).to(cuda_dev)
    B = _bf16(torch.randn(N, K, generator=g) / K ** 0.5).to(cuda_dev)
    bias = torch.randn(N, generator=g).to(cuda_dev)
    rbias = torch.randn(R, N, generator=g).to(cuda_dev)
    r1row = torch.randn(M, generator=g).to(cuda_dev)
    r1col = torch.randn(N, generator=g).to(cuda_dev)
    aux = _bf16(torch.randn(M, N, generator=g)).to(cuda_dev)
    hw = (torch.randn(3, N, generator=g) / N ** 0.5).to(cuda_dev)
    hb = torch.randn(3, generator=g).to(cuda_dev)
    hout = torch.full((M, 3), float("nan"), device=cuda_dev)
    Cout = torch.empty(M, N, dtype=torch.bfloat16, device=cuda_dev)
    ep = L.make_epilogue(bias=bias, ray_bias=rbias, rows_per_ray=S, rank1_row=r1row, rank1_col=r1col,
                         aux=aux if aux_mode else None, ldaux=N, aux_mode=aux_mode, act=1,
                         head_w=hw, head_b=hb, head_act=1, head_out=hout)
    L.gemm_bf16(A, B, Cout, M, N, K, ep=ep)
    torch.cuda.synchronize()
    v = A.float() @ B.float().t() + bias + rbias.repeat_interleave(S, 0) + r1row[:, None] * r1col
    if aux_mode == 1:
        v = v + aux.float()
    v = torch.relu(v)
    if aux_mode == 2:
        v = v * (aux.float() > 0)
    href = torch.nn.functional.softplus(v @ hw.t() + hb)
    assert torch.allclose(Cout.float(), v, rtol=1e-2, atol=2e-2)
    assert torch.allclose(hout, href, rtol=1e-3, atol=1e-3)


@pytest.mark.parametrize("M,N,K", [(4096, 256, 256), (64 * 1001, 256, 320), (8192, 128, 256),
                                   (12288, 256, 64), (64 * 77, 128, 128)])
def test_wgrad_bf16(cuda_dev, M, N, K):
    from upnerf_b200 import _lib as L

    g = torch.Generator(device="cpu").manual_seed(M + K)
    dY = _bf16(torch.randn(M, N, generator=g)).to(cuda_dev)
    X = _bf16(torch.randn(M, K, generator=g)).to(cuda_dev)
    # map packed columns [0,K-64) -> [63, ...) and the last 63 of the final 64 -> [0,63)
    if K > 64:
        segs = [(0, K - 64, 63), (K - 64, 63, 0)]
        kw = K - 1
    else:
        segs = [(0, 63, 0)]
        kw = 63
    dW = torch.zeros(N, kw, device=cuda_dev)
    db = torch.zeros(N, device=cuda_dev)
    L.wgrad_bf16(dY, X, dW, db, M, N, K, segs)
    torch.cuda.synchronize()
    full = dY.float().t() @ X.float()
    if K > 64:
        ref = torch.cat([full[:, K - 64:K - 1], full[:, :K - 64]], 1)
    else:
        ref = full[:, :63]
    scale = M ** 0.5
    assert torch.allclose(dW / scale, ref / scale, rtol=1e-3, atol=2e-3)
    assert torch.allclose(db / scale, dY.float().sum(0) / scale, rtol=1e-3, atol=2e-3)


@pytest.mark.parametrize("M,N,K", [(4096, 256, 256), (64 * 1001, 256, 320), (8192, 128, 256),
                                   (12288, 256, 64), (64 * 77, 128, 128), (64 * 3, 256, 256)])
def test_wgrad_bf16_deterministic(cuda_dev, M, N, K):
    """Atomic-free split reduction (the mode upnerf_render_bwd uses): same result as the atomic
    kernel within fp32 summation-order noise, accumulates into dW (+=), and is bit-reproducible."""
    from upnerf_b200 import _lib as L

    g = torch.Generator(device="cpu").manual_seed(M + K + 1)
    dY = _bf16(torch.randn(M, N, generator=g)).to(cuda_dev)
    X = _bf16(torch.randn(M, K, generator=g)).to(cuda_dev)
    if K > 64:
        segs, kw = [(0, K - 64, 63), (K - 64, 63, 0)], K - 1
    else:
        segs, kw = [(0, 63, 0)], 63
    init = torch.randn(N, kw, generator=g).to(cuda_dev)
    outs = []
    for _ in range(2):
        dW = init.clone()
        db = torch.ones(N, device=cuda_dev)
        L.wgrad_bf16_det(dY, X, dW, db, M, N, K, segs)
        torch.cuda.synchronize()
        outs.append((dW, db))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    full = dY.float().t() @ X.float()
    ref = torch.cat([full[:, K - 64:K - 1], full[:, :K - 64]], 1) if K > 64 else full[:, :63]
    scale = M ** 0.5
    assert torch.allclose((outs[0][0] - init) / scale, ref / scale, rtol=1e-3, atol=2e-3)
    assert torch.allclose((outs[0][1] - 1) / scale, dY.float().sum(0) / scale, rtol=1e-3, atol=2e-3)


@pytest.mark.parametrize("trans", ["nt", "nn", "tn"])
def test_gemm_f32_strided(cuda_dev, trans):
    from upnerf_b200 import _lib as L

    M, N, K = 333, 130, 77
    g = torch.Generator(device="cpu").manual_seed(5)
    if trans == "nt":   # C = A[M,K] B[N,K]^T
        A = torch.randn(M, K, generator=g).to(cuda_dev); B = torch.randn(N, K, generator=g).to(cuda_dev)
        sa, sb = (K, 1), (K, 1); ref = A @ B.t()
    elif trans == "nn":  # C = A[M,K] B[K,N]
        A = torch.randn(M, K, generator=g).to(cuda_dev); B = torch.randn(K, N, generator=g).to(cuda_dev)
        sa, sb = (K, 1), (1, N); ref = A @ B
    else:               # C = A[K,M]^T B[K,N]  (weight gradient form), split over k
        A = torch.randn(K, M, generator=g).to(cuda_dev); B = torch.randn(K, N, generator=g).to(cuda_dev)
        sa, sb = (1, M), (1, N); ref = A.t() @ B
    Cout = torch.zeros(M, N, device=cuda_dev)
    L.gemm_f32(A, sa, B, sb, Cout, (N, 1), M, N, K, split_k=3 if trans == "tn" else 1)
    torch.cuda.synchronize()
    assert torch.allclose(Cout, ref, rtol=1e-4, atol=1e-4)


def test_gemm_f32_epilogue(cuda_dev):
    from upnerf_b200 import _lib as L

    M, N, K, S = 256, 96, 40, 64
    g = torch.Generator(device="cpu").manual_seed(9)
    A = torch.randn(M, K, generator=g).to(cuda_dev)
    B = torch.randn(N, K, generator=g).to(cuda_dev)
    bias = torch.randn(N, generator=g).to(cuda_dev)
    rbias = torch.randn(M // S, N, generator=g).to(cuda_dev)
    aux = torch.randn(M, N, generator=g).to(cuda_dev)
    Cout = torch.ones(M, N, device=cuda_dev)
    ep = L.make_epilogue(bias=bias, ray_bias=rbias, rows_per_ray=S, aux=aux, ldaux=N, aux_mode=2, act=0)
    L.gemm_f32(A, (K, 1), B, (K, 1), Cout, (N, 1), M, N, K, ep=ep, accumulate=True)
    torch.cuda.synchronize()
    ref = (A @ B.t() + bias + rbias.repeat_interleave(S, 0)) * (aux > 0) + 1.0
    assert torch.allclose(Cout, ref, rtol=1e-4, atol=1e-4)


def _trunk_reference(pe, ws, bs, sw, sb):
    """Layer-by-layer fp32 math on the same bf16-rounded operands, activations rounded to bf16
    between layers exactly where the kernel rounds them (models/nerf.py:84-93)."""
    outs = []
    h = None
    x = pe.float()
    for l in range(9):
        if l == 0:
            a = x
        elif l == 4:
            a = torch.cat([h, x], 1)              # packed order [h | PE]
        else:
            a = h
        y = a @ ws[l].float().t() + bs[l]
        if l < 8:
            y = torch.relu(y)
        if l == 7:
            sig = torch.nn.functional.softplus(y @ sw + sb)
        h = y.to(torch.bfloat16).float()
        outs.append(h)
    return outs, sig


@pytest.mark.parametrize("variant", ["multicast", "single-cta"])
@pytest.mark.parametrize("M", [128, 128 * 3 + 37, 128 * 148 * 2 + 128 * 5 + 1, 128 * 148 * 5 + 77])
def test_mlp_trunk_fwd_fused(cuda_dev, M, variant, monkeypatch):
    """Fused trunk (PE -> 8 layers + skip -> final, sigma head) vs the layer-wise reference;
    covers a single tile, a ragged last tile, dummy tiles of an incomplete unit and several units per CTA pair
    (pipeline wrap-around), for CTA pairs sharing the weight stream by multicast (default) and single CTAs
    (UPNERF_TRUNK_CLUSTER=1)."""
    from upnerf_b200 import _lib as L

    monkeypatch.setenv("UPNERF_TRUNK_CLUSTER", "1" if variant == "single-cta" else "2")

    g = torch.Generator(device="cpu").manual_seed(M)
    pe = _bf16(torch.randn(M, 64, generator=g))
    pe[:, 63] = 0
    ks = [64, 256, 256, 256, 320, 256, 256, 256, 256]
    ws = [_bf16(torch.randn(256, k, generator=g) * (1.7 / k ** 0.5)) for k in ks]
    bs = [torch.randn(256, generator=g) * 0.1 for _ in ks]
    sw, sb = torch.randn(256, generator=g) / 16, torch.randn(1, generator=g)
    wcat = torch.cat(ws, 1).contiguous()
    assert wcat.shape == (256, L.TRUNK_WCAT_COLS)
    d = lambda t: t.to(cuda_dev)
    # outputs with different row strides: H4 lives in the 320-wide skip buffer
    outs = [torch.full((M, 320 if l == 3 else 256), float("nan"), dtype=torch.bfloat16, device=cuda_dev) for l in range(9)]
    sig = torch.full((M,), float("nan"), device=cuda_dev)
    L.mlp_trunk_fwd(d(pe), d(wcat), [d(b) for b in bs], d(sw), d(sb), [o[:, :256] for o in outs], sig, M)
    torch.cuda.synchronize()
    ref, ref_sig = _trunk_reference(d(pe), [d(w) for w in ws], [d(b) for b in bs], d(sw), d(sb))
    for l in range(9):
        got = outs[l][:, :256].float()
        assert torch.isfinite(got).all(), l
        # bf16 outputs of O(1) values: one rounding step (2^-8 relative) + re-association, growing
        # slowly with depth because each layer re-rounds its input
        err = (got - ref[l]).abs().max().item()
        scale = ref[l].abs().max().item()
        assert err <= 0.02 * max(scale, 1.0) * (1 + l / 4), (l, err, scale)
    assert torch.isnan(outs[3][:, 256:].float()).all()        # the PE columns of the skip buffer are untouched
    assert torch.allclose(sig, ref_sig, rtol=2e-2, atol=2e-2)


@pytest.mark.parametrize("variant", ["multicast", "single-cta"])
@pytest.mark.parametrize("M", [128 * 2 + 77, 128 * 148 * 2 + 128 * 3 + 9, 128 * 148 * 5 + 77])
def test_mlp_trunk_bwd_fused(cuda_dev, M, variant, monkeypatch):
    """Fused backward chain: dY8 = (dHF W_F + dssig (x) w_s) * [H8>0], dYl = (dY(l+1) W(l+1)) * [Hl>0],
    with the ReLU bit masks written by the fused forward, against layer-wise fp32 math on the same
    bf16 operands (masks taken from the forward kernel's own activations, so no mask can flip); CTA pairs and
    single CTAs (UPNERF_TRUNK_CLUSTER=1)."""
    from upnerf_b200 import _lib as L

    monkeypatch.setenv("UPNERF_TRUNK_CLUSTER", "1" if variant == "single-cta" else "2")
    g = torch.Generator(device="cpu").manual_seed(M + 1)
    d = lambda t: t.to(cuda_dev)
    pe = _bf16(torch.randn(M, 64, generator=g))
    pe[:, 63] = 0
    ks = [64, 256, 256, 256, 320, 256, 256, 256, 256]
    ws = [d(_bf16(torch.randn(256, k, generator=g) * (1.7 / k ** 0.5))) for k in ks]
    bs = [d(torch.randn(256, generator=g) * 0.1) for _ in ks]
    sw, sb = d(torch.randn(256, generator=g) / 16), d(torch.randn(1, generator=g))
    wcat = torch.cat(ws, 1).contiguous()
    outs = [torch.empty(M, 256, dtype=torch.bfloat16, device=cuda_dev) for _ in ks]
    sig = torch.empty(M, device=cuda_dev)
    mask = torch.zeros(L.trunk_mask_words(M), dtype=torch.int32, device=cuda_dev)
    L.mlp_trunk_fwd(d(pe), wcat, bs, sw, sb, outs, sig, M, relu_mask=mask)
    # transposed weights in chain order: WF | W8 | W7 | W6 | W5[h part] | W4 | W3 | W2, each [in, out]
    chain = [ws[8], ws[7], ws[6], ws[5], ws[4][:, :256], ws[3], ws[2], ws[1]]
    wcat_t = torch.cat([w.t().contiguous() for w in chain], 1).contiguous()
    assert wcat_t.shape == (256, L.TRUNK_WCATT_COLS)
    d_hf = d(_bf16(torch.randn(M, 256, generator=g)))
    d_ssig = d(torch.randn(M, generator=g))
    d_outs = [torch.full((M, 256), float("nan"), dtype=torch.bfloat16, device=cuda_dev) for _ in range(8)]
    L.mlp_trunk_bwd(d_hf, d_ssig, sw, wcat_t, mask, d_outs, M)
    torch.cuda.synchronize()
    cur = d_hf.float()
    for j in range(8):
        y = cur @ chain[j].float()                      # [M, out] @ [out, in]
        if j == 0:
            y = y + d_ssig[:, None] * sw[None, :]
        y = y * (outs[7 - j].float() > 0)
        got = d_outs[j].float()
        assert torch.isfinite(got).all(), j
        err = (got - y).abs().max().item()
        assert err <= 0.02 * max(y.abs().max().item(), 1.0), (j, err)
        assert ((got != 0) <= (outs[7 - j].float() > 0)).all(), j      # the mask is exact
        cur = got                                        # chain on the kernel's own bf16 output


def _tf32_round(x):
    """Round-to-nearest (ties away) to tf32's 10-bit mantissa, as cvt.rna.tf32.f32 does."""
    i = x.contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32)


@pytest.mark.parametrize("case", ["nt", "nn", "tn_split", "tn_big", "vec_m1", "outer_k1", "n16", "col_major_out"])
def test_gemm_tf32_strided(cuda_dev, case):
    """upnerf_gemm_tf32 (tcgen05 kind::tf32, thread-staged operands of any stride) against an fp64 product of the
    tf32-rounded operands (tight) and of the raw operands (tf32 resolution) -- the shapes and stride patterns of
    the per-ray / parameter-space products in render.cu."""
    from upnerf_b200 import _lib as L

    g = torch.Generator(device="cpu").manual_seed(21)
    rnd = lambda *s: torch.randn(*s, generator=g).to(cuda_dev)
    split, acc, ep, sc = 1, False, None, None
    if case == "nt":            # per-ray bias: C[R,128] = P[R,75] W[128, cols]^T + b   (W a column slice: ld > K)
        M, N, K = 4096 + 37, 128, 75
        A = rnd(M, K); Wfull = rnd(N, 459); B = Wfull[:, 384:]
        sa, sb = (K, 1), (459, 1); ref_ops = (A, B.t())
        bias = rnd(N); ep = L.make_epilogue(bias=bias)
    elif case == "nn":          # data gradient: gHFr[R,256] = gf[R,384] Wsf[384,256]
        M, N, K = 1000, 256, 384
        A = rnd(M, K); B = rnd(K, N)
        sa, sb = (K, 1), (1, N); ref_ops = (A, B)
    elif case == "tn_split":    # weight gradient over rays with atomics: dW[384,256] += gf[R,384]^T HFr[R,256]
        M, N, K = 384, 256, 4096
        A = rnd(K, M); B = rnd(K, N)
        sa, sb = (1, M), (1, N); ref_ops = (A.t(), B); split = 32
    elif case == "tn_big":      # odd sizes, K not a multiple of the split or of 32
        M, N, K = 130, 77, 1001
        A = rnd(K, M); B = rnd(K, N)
        sa, sb = (1, M), (1, N); ref_ops = (A.t(), B); split = 7
    elif case == "vec_m1":      # bq_const[128] = b_sf[384] W[128,384]^T + b   (M = 1, broadcast row stride)
        M, N, K = 1, 128, 384
        A = rnd(1, K); B = rnd(N, K)
        sa, sb = (0, 1), (K, 1); ref_ops = (A, B.t())
        bias = rnd(N); ep = L.make_epilogue(bias=bias); acc = True
    elif case == "outer_k1":    # rank-1 update, K = 1, accumulate
        M, N, K = 128, 384, 1
        A = rnd(M, 1); B = rnd(N, 1)
        sa, sb = (1, 0), (1, 0); ref_ops = (A, B.t()); acc = True
    elif case == "n16":         # dCrows[R,16] = dBc[R,128] W[128, 256:272]  (narrow output, strided B)
        M, N, K = 4096, 16, 128
        A = rnd(M, K); Wfull = rnd(K, 272); B = Wfull[:, 256:]
        sa, sb = (K, 1), (1, 272); ref_ops = (A, B)
    else:                       # transposed (column-major) output + rank-1 term
        M, N, K = 200, 300, 64
        A = rnd(M, K); B = rnd(N, K)
        sa, sb = (K, 1), (K, 1); ref_ops = (A, B.t())
        r1r, r1c = rnd(M), rnd(N); ep = L.make_epilogue(rank1_row=r1r, rank1_col=r1c)
        sc = (1, M)
    init = rnd(N, M).t() if sc else rnd(M, N)
    Cout = init.clone() if (acc or split > 1) else torch.full_like(init, 7.0)
    if sc is None:
        sc = (N, 1)
    else:
        Cout = Cout.t().contiguous().t()          # [M,N] view of a column-major buffer
        init = Cout.clone()
    L.gemm_tf32(A, sa, B, sb, Cout, sc, M, N, K, ep=ep, accumulate=acc, split_k=split)
    torch.cuda.synchronize()
    extra = torch.zeros(M, N, device=cuda_dev, dtype=torch.float64)
    if case in ("nt", "vec_m1"):
        extra = extra + bias.double()
    if case == "col_major_out":
        extra = extra + r1r.double()[:, None] * r1c.double()[None, :]
    if acc or split > 1:
        extra = extra + init.double()
    a64, b64 = ref_ops[0].double(), ref_ops[1].double()
    ref_exact = a64 @ b64 + extra
    ref_tf32 = _tf32_round(ref_ops[0].contiguous()).double() @ _tf32_round(ref_ops[1].contiguous()).double() + extra
    scale = float((a64.abs() @ b64.abs()).max()) + 1.0
    err_t = float((Cout.double() - ref_tf32).abs().max()) / scale
    err_e = float((Cout.double() - ref_exact).abs().max()) / scale
    print(f"{case}: vs tf32-rounded operands {err_t:.2e}, vs exact {err_e:.2e}")
    assert err_t <= 2e-6, err_t          # fp32 accumulation of exact tf32 products
    assert err_e <= 1e-3, err_e          # tf32 operand resolution (2^-11 per operand)


@pytest.mark.parametrize("M", [128 * 5 + 3, 128 * 148 * 2 + 77])
def test_gemm2_bf16_two_operands(cuda_dev, M):
    """upnerf_gemm2_bf16: C = [A1 | A2] B^T in one pass (the positional-encoding gradient dPE = dY5 W5[:, pe] + dY1 W1)
    against fp32 math on the same bf16 operands."""
    from upnerf_b200 import _lib as L

    g = torch.Generator(device="cpu").manual_seed(M)
    A1 = _bf16(torch.randn(M, 256, generator=g)).to(cuda_dev)
    A2 = _bf16(torch.randn(M, 256, generator=g)).to(cuda_dev)
    B = _bf16(torch.randn(64, 512, generator=g) / 16).to(cuda_dev)
    Cc = torch.full((M, 64), float("nan"), dtype=torch.bfloat16, device=cuda_dev)
    L.gemm2_bf16(A1, A2, B, Cc, M, 64, 256, 256)
    torch.cuda.synchronize()
    ref = A1.float() @ B[:, :256].float().t() + A2.float() @ B[:, 256:].float().t()
    err = float((Cc.float() - ref).abs().max())
    assert torch.isfinite(Cc.float()).all()
    assert err <= 0.02 * max(1.0, float(ref.abs().max())), err
