// Host-side chaining of the generic MLP kernels (cstr_mlp.cuh) with the shared GEMM / skinny / finalize kernels of cstr_td3.cu.
// Included inside cstr_td3.cu's anonymous namespace, after forward_hidden / backward_hidden, whose structure these follow.
#pragma once

// bump allocator over the caller's workspace (base == nullptr: size query)
struct Bump {
    float *base;
    int64_t o = 0;
    explicit Bump(float *b) : base(b) {}
    float *take(int64_t n) {
        float *p = base ? base + o : nullptr;
        o += pad4(n);
        return p;
    }
};

// scratch the shared backward kernels need (what `Workspace` carries for TD3), sized for hidden widths up to HM and Z nets per call
struct MlpScratch {
    float *slabs, *skinny;
    int64_t slab_cap, skinny_region;
    int tensor;
};

inline MlpScratch mlp_scratch(Bump &b, int rows, int HM, int Z, int max_skinny_ny) {
    MlpScratch s;
    s.slab_cap = (int64_t)MAX_SPLITS * Z * pad4((int64_t)HM * HM);
    s.slabs = b.take(s.slab_cap);
    const int64_t hp = ((int64_t)HM + 31) / 32 * 32;
    s.skinny_region = (int64_t)((rows + SKINNY_ROWS - 1) / SKINNY_ROWS) * Z * (max_skinny_ny + 1) * hp;
    s.skinny = b.take(3 * s.skinny_region);
    s.tensor = 0;
    return s;
}

inline Workspace as_workspace(const MlpScratch &s) {  // the fields backward_hidden-style code reads
    Workspace w{};
    w.slabs = s.slabs, w.slab_cap = s.slab_cap, w.skinny = s.skinny, w.skinny_region = s.skinny_region, w.tensor = s.tensor;
    return w;
}

// h1 = relu(L1(x)), h2 = relu(L2(h1)) for Z nets `z_stride` floats apart (activations: Z slabs of B rows)
int forward_hidden_g(int B, const NetLayout &L, const Src &s, const Net &n, int64_t z_stride, int Z, float *h1, float *h2, const MlpScratch &sc,
                     cudaStream_t st) {
    if (s.n0 + s.n1 != L.in) return fail_arg(CSTR_EINVAL, "mlp: source widths do not add up to the net's input width");
    const int64_t threads = (int64_t)((B + L1G_ROWS - 1) / L1G_ROWS) * (L.h1 / 4);
    launch_k(mlp_layer1_kernel, dim3((unsigned)((threads + 255) / 256), Z), 256, 0, st, B, L.h1, s, (const float *)n.w1, (const float *)n.b1, z_stride, h1,
             (int64_t)B * L.h1);
    if (int rc = check_launch("mlp_layer1_kernel")) return rc;
    GemmArgs g{};
    g.A = h1, g.Bm = n.w2, g.aux = n.b2, g.C = h2;
    g.M = B, g.N = L.h2, g.K = L.h1, g.lda = L.h1, g.ldb = L.h1, g.ldc = L.h2, g.ldaux = 0;
    g.a_z = (int64_t)B * L.h1, g.b_z = z_stride, g.c_z = (int64_t)B * L.h2, g.aux_z = z_stride;
    g.splits = 1, g.k_per_split = L.h1, g.c_split = 0;
    g.split_buf = sc.slabs, g.split_cap = sc.slab_cap;
    return launch_gemm<G_FWD>(g, Z, sc.tensor, st, "td3_gemm_kernel<fwd>");
}

// ... followed by the linear head: y (Z, B, out), tanh when squash
int forward_mlp(int B, const NetLayout &L, const Src &s, const Net &n, int64_t z_stride, int Z, float *h1, float *h2, float *y, bool squash,
                const MlpScratch &sc, cudaStream_t st) {
    if (int rc = forward_hidden_g(B, L, s, n, z_stride, Z, h1, h2, sc, st)) return rc;
    launch_k(mlp_head_fwd_kernel, dim3((B + 7) / 8, Z), 256, 0, st, B, L.h2, L.out, (const float *)h2, (int64_t)B * L.h2, (const float *)n.w3, (const float *)n.b3,
             z_stride, y, squash ? 1 : 0);
    return check_launch("mlp_head_fwd_kernel");
}

template <bool YBIAS>
int launch_skinny_ny(int ny, SkinnyArgs s, int Z, float *part, cudaStream_t st, const char *what, FinJobs *defer) {
    switch (ny) {
        case 0: return launch_skinny<0, false>(s, Z, part, st, what, defer);
        case 1: return launch_skinny<1, YBIAS>(s, Z, part, st, what, defer);
        case 2: return launch_skinny<2, YBIAS>(s, Z, part, st, what, defer);
        case 3: return launch_skinny<3, YBIAS>(s, Z, part, st, what, defer);
        case 4: return launch_skinny<4, YBIAS>(s, Z, part, st, what, defer);
        case 6: return launch_skinny<6, YBIAS>(s, Z, part, st, what, defer);
        default: return fail_arg(CSTR_EINVAL, "mlp: skinny weight gradient supports widths 0-4 and 6");
    }
}

// plain FFMA C (M,N) = A^T-role GEMM over the batch, written straight into the gradient tensor (one split: fixed summation order)
int wgrad_gemm_direct(int M, int N, int K, const float *A, int lda, const float *Bm, int ldb, float *C, int ldc, cudaStream_t st, const char *what) {
    if ((M & 3) || (N & 3) || (lda & 3) || (ldb & 3) || (ldc & 3)) return fail_arg(CSTR_EINVAL, "mlp: wide weight gradients need widths that are multiples of 4");
    GemmArgs q{};
    q.A = A, q.Bm = Bm, q.aux = nullptr, q.C = C;
    q.M = M, q.N = N, q.K = K, q.lda = lda, q.ldb = ldb, q.ldc = ldc, q.ldaux = 0;
    q.splits = 1, q.k_per_split = K, q.c_split = 0;
    return launch_gemm<G_WGRAD>(q, 1, 0, st, what);
}

// Given dy (B, out; tanh' already applied by the caller) and the forward activations: every weight gradient of ONE net into `gn`, and dz1 left in
// `dz1` for a following mlp_dx_kernel.  x_wide: the contiguous (B, in) input, needed (instead of the skinny kernel) when in > 6.
int backward_mlp(int B, const NetLayout &L, const Src &s, const float *x_wide, const Net &n, const Net &gn, const float *h1, const float *h2,
                 const float *dy, float *dz2, float *dz1, const MlpScratch &sc, bool want_weight_grads, cudaStream_t st) {
    const int H1 = L.h1, H2 = L.h2, OUT = L.out;
    launch_k(mlp_head_bwd_kernel, (B + 7) / 8, 256, 0, st, B, H2, OUT, dy, (const float *)n.w3, h2, dz2);
    if (int rc = check_launch("mlp_head_bwd_kernel")) return rc;
    FinJobs J{};
    if (want_weight_grads) {
        if (OUT <= 4) {  // dW3 = dy^T @ h2, db3 = sum_b dy
            SkinnyArgs k{};
            k.X = h2, k.x_z = 0, k.ldx = H2, k.H = H2, k.B = B;
            k.Y0 = dy, k.n0 = OUT, k.ld0 = OUT, k.Y1 = nullptr, k.n1 = 0, k.ld1 = 0, k.y_z = 0;
            k.out_w = gn.w3, k.out_b = gn.b3, k.out_z = 0, k.transposed = 1;
            if (int rc = launch_skinny_ny<true>(OUT, k, 1, sc.skinny + 2 * sc.skinny_region, st, "td3_skinny_wgrad_kernel<w3>", &J)) return rc;
        } else {  // a wide head (the VAE's [mean; log_std]): GEMM over the batch + a column sum for the bias
            if (int rc = wgrad_gemm_direct(OUT, H2, B, dy, OUT, h2, H2, gn.w3, H2, st, "td3_gemm_kernel<wgrad head>")) return rc;
            SkinnyArgs k{};
            k.X = dy, k.x_z = 0, k.ldx = OUT, k.H = OUT, k.B = B;
            k.out_w = nullptr, k.out_b = gn.b3, k.out_z = 0;
            if (int rc = launch_skinny<0, false>(k, 1, sc.skinny + 2 * sc.skinny_region, st, "td3_skinny_wgrad_kernel<b3>", &J)) return rc;
        }
    }
    GemmArgs g{};  // dz1 = (dz2 @ W2) * (h1 > 0)
    g.A = dz2, g.Bm = n.w2, g.aux = h1, g.C = dz1;
    g.M = B, g.N = H1, g.K = H2, g.lda = H2, g.ldb = H1, g.ldc = H1, g.ldaux = H1;
    g.splits = 1, g.k_per_split = H2, g.c_split = 0;
    g.split_buf = sc.slabs, g.split_cap = sc.slab_cap;
    if (int rc = launch_gemm<G_DGRAD>(g, 1, sc.tensor, st, "td3_gemm_kernel<dgrad>")) return rc;
    if (!want_weight_grads) return 0;
    const int64_t w2n = pad4((int64_t)H1 * H2);  // dW2 = dz2^T @ h1, split over the batch into slabs, summed in order by the finalize launch
    GemmArgs q{};
    q.A = dz2, q.Bm = h1, q.aux = nullptr, q.C = sc.slabs;
    q.M = H2, q.N = H1, q.K = B, q.lda = H2, q.ldb = H1, q.ldc = H1, q.ldaux = 0;
    q.c_z = w2n;
    const int splits = choose_splits(B, H1, H2, 1, sc.tensor);
    q.splits = splits, q.k_per_split = (B + splits - 1) / splits, q.c_split = w2n;
    if (int rc = launch_gemm<G_WGRAD>(q, 1, sc.tensor, st, "td3_gemm_kernel<wgrad>")) return rc;
    J.slab_n4 = (int64_t)H1 * H2 / 4, J.slab_splits = splits, J.slabs = (const float4 *)sc.slabs, J.slab_stride4 = w2n / 4, J.slab_z4 = w2n / 4;
    J.slab_out = (float4 *)gn.w2, J.slab_out_z4 = 0, J.slab_Z = 1;
    SkinnyArgs b2{};  // db2 = colsum(dz2)
    b2.X = dz2, b2.x_z = 0, b2.ldx = H2, b2.H = H2, b2.B = B;
    b2.out_w = nullptr, b2.out_b = gn.b2, b2.out_z = 0;
    if (int rc = launch_skinny<0, false>(b2, 1, sc.skinny, st, "td3_skinny_wgrad_kernel<b2>", &J)) return rc;
    if (L.in <= 6 && L.in != 5) {  // dW1 = dz1^T @ x, db1 = colsum(dz1)
        if (s.x0_rows) return fail_arg(CSTR_EINVAL, "mlp: weight gradients over a repeated source are not supported");
        SkinnyArgs t{};
        t.X = dz1, t.x_z = 0, t.ldx = H1, t.H = H1, t.B = B;
        t.Y0 = s.x0, t.n0 = s.n0, t.ld0 = s.ld0, t.Y1 = s.x1, t.n1 = s.n1, t.ld1 = s.ld1, t.y_z = 0;
        t.out_w = gn.w1, t.out_b = gn.b1, t.out_z = 0;
        if (int rc = launch_skinny_ny<false>(L.in, t, 1, sc.skinny + sc.skinny_region, st, "td3_skinny_wgrad_kernel<w1>", &J)) return rc;
    } else {
        if (!x_wide) return fail_arg(CSTR_EINVAL, "mlp: a layer-1 input wider than 6 needs its contiguous copy for the weight gradient");
        if (int rc = wgrad_gemm_direct(H1, L.in, B, dz1, H1, x_wide, L.in, gn.w1, L.in, st, "td3_gemm_kernel<wgrad layer1>")) return rc;
        SkinnyArgs t{};
        t.X = dz1, t.x_z = 0, t.ldx = H1, t.H = H1, t.B = B;
        t.out_w = nullptr, t.out_b = gn.b1, t.out_z = 0;
        if (int rc = launch_skinny<0, false>(t, 1, sc.skinny + sc.skinny_region, st, "td3_skinny_wgrad_kernel<b1>", &J)) return rc;
    }
    return launch_finalize(J, st);
}
