/*
 * sdempc_oracle_impl.h — CPU restatement (ORACLE) of the neural-SDE MPC solve.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path may include, link or
 * call this file; only tests/, __graft_entry__.smoke() and bench.py's CPU
 * baseline legs use it, as the checker / the timed CPU baseline.
 *
 * PARITY UNPINNED: the arithmetic of the reference solve lives in the
 * un-vendored, un-pinned package github.com/wuwushrek/sde4mbrl (imported at
 * /root/reference sde4mbrl_px4/mpc_controller/sde_control.py:12-13, built at
 * :685, called at :345-350 and :400-416).  Its source, its JAX runtime, its
 * model pickles and any golden vectors are absent, so this file restates the
 * algorithm from the reference's call sites, its YAML schema
 * (launch/iris_sitl_traj_mpc.yaml:1-85) and the canonical [SPEC] block of
 * SURVEY.md section 8(a).  The only externally pinned piece is the Philox4x32-10
 * generator (Random123 known-answer vectors, tests/test_oracle.py).
 *
 * The file is a template over REAL (float / double); sdempc_oracle.c includes
 * it twice.  It is written as plain scalar loops over units, particles and
 * steps — unlike the warp-distributed CUDA kernels it checks — but follows the
 * operation ORDER of DESIGN.md "Arithmetic specification" (SPEC-ARITH).
 */

#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(CAT(oracle_, SUFFIX), CAT(_, name))

/* SPEC-ARITH: the float32 instance is a fixed sequence of IEEE add/mul/fma
 * (explicit FMA() calls, compiled with -ffp-contract=off) plus the det_* elementary
 * functions of det_math.h, so that the CUDA kernels can reproduce it bit for bit.
 * The float64 instance keeps the same expression trees but uses libm; it exists
 * for finite-difference and accuracy checks of the float32 one. */
#ifdef REAL_IS_DOUBLE
#define FMA(a, b, c) fma((a), (b), (c))
#define R_SQRT sqrt
#define R_FABS fabs
#define M_TANH(x) tanh(x)
#define M_SOFTPLUS(s) (((s) > 0 ? (s) : 0) + log1p(exp(-fabs(s))))
#define M_SIGMOID(s) (1.0 / (1.0 + exp(-(s))))
#define M_RSQRT1(n2) (1.0 / sqrt(n2))
#define M_LOG(u) log(u)
#define M_SINCOS2PI(u, s, c) do { *(s) = sin(6.28318530717958647692 * (u)); *(c) = cos(6.28318530717958647692 * (u)); } while (0)
#else
#define FMA(a, b, c) fmaf((a), (b), (c))
#define R_SQRT sqrtf
#define R_FABS fabsf
#define M_TANH(x) det_tanh(x)
#define M_SOFTPLUS(s) det_softplus(s)
#define M_SIGMOID(s) det_sigmoid(s)
#define M_RSQRT1(n2) det_rsqrt_near1(n2)
#define M_LOG(u) det_log(u)
#define M_SINCOS2PI(u, s, c) det_sincos2pi((u), (s), (c))
#endif

#define NX SDEMPC_NX
#define MAXW 128
#define NUH (SDEMPC_MAX_H * SDEMPC_MAX_NU)

/* ---- model view ----------------------------------------------------------- */
typedef struct {
    int nu, n_in, width;
    REAL inv_m, gravity, kT, kT2, J[3], Jinv[3], Jd[3], mixer[3][SDEMPC_MAX_NU], sig0[6];
    /* [net 0 = drift, 1 = diffusion] */
    const float *W1[2], *b1[2], *W2[2], *b2[2], *W3[2], *b3[2];
    float *W1T[2], *W2T[2];   /* transposed copies [in][out] for the vectorised forward layers */
} CAT(omodel_, SUFFIX);
#define OMODEL CAT(omodel_, SUFFIX)

static int FN(parse_model)(const void* blob, size_t nbytes, OMODEL* m) {
    if (nbytes < sizeof(sdempc_model_header)) return -1;
    const sdempc_model_header* h = (const sdempc_model_header*)blob;
    if (h->magic != SDEMPC_MODEL_MAGIC || h->version != SDEMPC_MODEL_VERSION) return -1;
    if (h->n_hidden != 2 || h->n_out != 6 || h->n_in != 6 + h->nu) return -1;
    if (h->width < 1 || h->width > MAXW || h->nu < 1 || h->nu > SDEMPC_MAX_NU) return -1;
    int W = h->width, n_in = h->n_in;
    size_t per_net = (size_t)W * n_in + W + (size_t)W * W + W + 6 * (size_t)W + 6;
    if (nbytes < sizeof(*h) + 2 * per_net * sizeof(float)) return -1;
    m->nu = h->nu; m->n_in = n_in; m->width = W;
    /* derived constants: one IEEE operation each, shared with the CUDA library's host code */
    m->inv_m = (REAL)1 / (REAL)h->mass; m->gravity = h->gravity; m->kT = h->k_thrust;
    m->kT2 = (REAL)2 * (REAL)h->k_thrust;
    for (int i = 0; i < 3; ++i) { m->J[i] = h->inertia[i]; m->Jinv[i] = (REAL)1 / (REAL)h->inertia[i]; }
    m->Jd[0] = (REAL)h->inertia[2] - (REAL)h->inertia[1];
    m->Jd[1] = (REAL)h->inertia[0] - (REAL)h->inertia[2];
    m->Jd[2] = (REAL)h->inertia[1] - (REAL)h->inertia[0];
    for (int r = 0; r < 3; ++r)
        for (int i = 0; i < SDEMPC_MAX_NU; ++i) m->mixer[r][i] = h->mixer[r * SDEMPC_MAX_NU + i];
    for (int i = 0; i < 6; ++i) m->sig0[i] = h->sigma_prior[i];
    const float* p = (const float*)((const char*)blob + sizeof(*h));
    for (int n = 0; n < 2; ++n) {
        m->W1[n] = p; p += (size_t)W * n_in;
        m->b1[n] = p; p += W;
        m->W2[n] = p; p += (size_t)W * W;
        m->b2[n] = p; p += W;
        m->W3[n] = p; p += 6 * (size_t)W;
        m->b3[n] = p; p += 6;
        m->W1T[n] = (float*)malloc(sizeof(float) * (size_t)W * n_in);
        m->W2T[n] = (float*)malloc(sizeof(float) * (size_t)W * W);
        for (int j = 0; j < W; ++j) {
            for (int k = 0; k < n_in; ++k) m->W1T[n][(size_t)k * W + j] = m->W1[n][(size_t)j * n_in + k];
            for (int k = 0; k < W; ++k) m->W2T[n][(size_t)k * W + j] = m->W2[n][(size_t)j * W + k];
        }
    }
    return 0;
}

/* ---- frames ([SPEC] "State / frames") ------------------------------------- */
/* enu2ned(x, np) of the reference (sde_control.py:13, 400):
 * p,v: (x,y,z)->(y,x,-z); w: (x,y,z)->(x,-y,-z);
 * q_ned = q_r (x) q_enu (x) q_b, q_r = (0,s,s,0), q_b = (0,1,0,0), s = sqrt(1/2),
 * which expands to +-s*(w+z, x+y, x-y, w-z), sign chosen so that qw >= 0.
 * The map is an involution, so the same routine serves as ned2enu. */
static void FN(enu_ned)(const REAL x[NX], REAL o[NX]) {
    const REAL s = (REAL)0.70710678118654752440;
    REAL c0 = s * (x[6] + x[9]), c1 = s * (x[7] + x[8]), c2 = s * (x[7] - x[8]), c3 = s * (x[6] - x[9]);
    if (c0 < 0) { c0 = -c0; c1 = -c1; c2 = -c2; c3 = -c3; }
    REAL t0 = x[0], t3 = x[3];
    o[0] = x[1]; o[1] = t0; o[2] = -x[2];
    o[3] = x[4]; o[4] = t3; o[5] = -x[5];
    o[6] = c0; o[7] = c1; o[8] = c2; o[9] = c3;
    o[10] = x[10]; o[11] = -x[11]; o[12] = -x[12];
}

/* ---- Philox4x32-10 + Box-Muller ([SPEC] "Noise") -------------------------- */
#ifndef ORACLE_PHILOX_DEFINED
#define ORACLE_PHILOX_DEFINED
static void oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
/* counter = (t, p, tick_lo, (tick_hi & 0x3fffffff) | sub << 30), key = seed */
static void oracle_noise_block(uint64_t seed, uint64_t tick, uint32_t t, uint32_t p, uint32_t sub, uint32_t out[4]) {
    uint32_t ctr[4] = {t, p, (uint32_t)tick, (((uint32_t)(tick >> 32)) & 0x3FFFFFFFu) | (sub << 30)};
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    oracle_philox4x32_10(ctr, key, out);
}
void oracle_philox(const uint32_t* ctr, const uint32_t* key, uint32_t* out) { oracle_philox4x32_10(ctr, key, out); }
#endif

static void FN(box_muller)(uint32_t a, uint32_t b, REAL* n0, REAL* n1) {
    /* uniforms in (0,1] from the top 24 bits, float32 arithmetic: for a >> 8 >= 2^23 the half-cell offset
     * rounds to even (u = 1 is reachable, u = 0 is not), tests/test_independent_apg.py restates this */
    float u1 = ((float)(a >> 8) + 0.5f) * 5.9604644775390625e-08f;
    float u2 = ((float)(b >> 8) + 0.5f) * 5.9604644775390625e-08f;
    REAL rad = R_SQRT((REAL)-2 * M_LOG((REAL)u1));
    REAL sn, cs;
    M_SINCOS2PI((REAL)u2, &sn, &cs);
    *n0 = rad * cs;
    *n1 = rad * sn;
}

/* xi[P][H][6], sub-stream `sub0` (0 for the solver, 2 for the plant) */
static void FN(gen_noise)(uint64_t seed, uint64_t tick, int P, int H, uint32_t sub0, REAL* xi) {
    for (int p = 0; p < P; ++p)
        for (int t = 0; t < H; ++t) {
            uint32_t r[4];
            REAL* o = xi + ((size_t)p * H + t) * 6;
            oracle_noise_block(seed, tick, (uint32_t)t, (uint32_t)p, sub0, r);
            FN(box_muller)(r[0], r[1], &o[0], &o[1]);
            FN(box_muller)(r[2], r[3], &o[2], &o[3]);
            oracle_noise_block(seed, tick, (uint32_t)t, (uint32_t)p, sub0 + 1, r);
            FN(box_muller)(r[0], r[1], &o[4], &o[5]);
        }
}

/* ---- SPEC-ARITH dot product: four partial sums over k mod 4, bias in lane 0 ---- */
static REAL FN(dot4)(int n, const float* w, int wstride, const REAL* v, REAL bias) {
    REAL a0 = bias, a1 = 0, a2 = 0, a3 = 0;
    int k = 0;
    for (; k + 3 < n; k += 4) {
        a0 = FMA((REAL)w[(size_t)k * wstride], v[k], a0);
        a1 = FMA((REAL)w[(size_t)(k + 1) * wstride], v[k + 1], a1);
        a2 = FMA((REAL)w[(size_t)(k + 2) * wstride], v[k + 2], a2);
        a3 = FMA((REAL)w[(size_t)(k + 3) * wstride], v[k + 3], a3);
    }
    if (k < n) a0 = FMA((REAL)w[(size_t)k * wstride], v[k], a0);
    if (k + 1 < n) a1 = FMA((REAL)w[(size_t)(k + 1) * wstride], v[k + 1], a1);
    if (k + 2 < n) a2 = FMA((REAL)w[(size_t)(k + 2) * wstride], v[k + 2], a2);
    return (a0 + a1) + (a2 + a3);
}

/* ---- MLP ------------------------------------------------------------------- */
typedef struct {
    REAL z[6 + SDEMPC_MAX_NU];
    REAL h1[2][MAXW], h2[2][MAXW], out[2][6];
} CAT(otape_, SUFFIX);
#define OTAPE CAT(otape_, SUFFIX)

/* The layer loops run over the OUTPUT index in the inner loop (contiguous, auto-vectorised by gcc) while each
 * output keeps the SPEC-ARITH accumulation order of dot4: partial sum (k mod 4), bias in partial 0, combined
 * as (a0+a1)+(a2+a3).  Forward layers 1-2 use the transposed copies W1T[k][j], W2T[k][j] made at create(). */
static void FN(mlp_fwd)(const OMODEL* m, OTAPE* tp) {
    const int W = m->width, n_in = m->n_in;
    REAL acc[4][MAXW];
    for (int n = 0; n < 2; ++n) {
        for (int j = 0; j < W; ++j) { acc[0][j] = (REAL)m->b1[n][j]; acc[1][j] = 0; acc[2][j] = 0; acc[3][j] = 0; }
        for (int k = 0; k < n_in; ++k) {
            const float* w = m->W1T[n] + (size_t)k * W;
            REAL* a = acc[k & 3];
            const REAL zk = tp->z[k];
            for (int j = 0; j < W; ++j) a[j] = FMA((REAL)w[j], zk, a[j]);
        }
        for (int j = 0; j < W; ++j) tp->h1[n][j] = M_TANH((acc[0][j] + acc[1][j]) + (acc[2][j] + acc[3][j]));
        for (int j = 0; j < W; ++j) { acc[0][j] = (REAL)m->b2[n][j]; acc[1][j] = 0; acc[2][j] = 0; acc[3][j] = 0; }
        for (int k = 0; k < W; ++k) {
            const float* w = m->W2T[n] + (size_t)k * W;
            REAL* a = acc[k & 3];
            const REAL hk = tp->h1[n][k];
            for (int j = 0; j < W; ++j) a[j] = FMA((REAL)w[j], hk, a[j]);
        }
        for (int j = 0; j < W; ++j) tp->h2[n][j] = M_TANH((acc[0][j] + acc[1][j]) + (acc[2][j] + acc[3][j]));
        for (int o = 0; o < 6; ++o)
            tp->out[n][o] = FN(dot4)(W, m->W3[n] + (size_t)o * W, 1, tp->h2[n], (REAL)m->b3[n][o]);
    }
}

/* lz[n_in] = J_drift^T lout[0] + J_diff^T lout[1] */
static void FN(mlp_bwd)(const OMODEL* m, const OTAPE* tp, REAL lout[2][6], REAL* lz) {
    const int W = m->width, n_in = m->n_in;
    REAL acc[4][MAXW], d2[MAXW], d1[MAXW], lzn[2][6 + SDEMPC_MAX_NU];
    for (int n = 0; n < 2; ++n) {
        for (int c4 = 0; c4 < 4; ++c4) for (int j = 0; j < W; ++j) acc[c4][j] = 0;
        for (int o = 0; o < 6; ++o) {
            const float* w = m->W3[n] + (size_t)o * W;
            REAL* a = acc[o & 3];
            const REAL lo = lout[n][o];
            for (int j = 0; j < W; ++j) a[j] = FMA((REAL)w[j], lo, a[j]);
        }
        for (int j = 0; j < W; ++j)
            d2[j] = ((acc[0][j] + acc[1][j]) + (acc[2][j] + acc[3][j])) * FMA(-tp->h2[n][j], tp->h2[n][j], (REAL)1);
        for (int c4 = 0; c4 < 4; ++c4) for (int k = 0; k < W; ++k) acc[c4][k] = 0;
        for (int j = 0; j < W; ++j) {
            const float* w = m->W2[n] + (size_t)j * W;
            REAL* a = acc[j & 3];
            const REAL dj = d2[j];
            for (int k = 0; k < W; ++k) a[k] = FMA((REAL)w[k], dj, a[k]);
        }
        for (int k = 0; k < W; ++k)
            d1[k] = ((acc[0][k] + acc[1][k]) + (acc[2][k] + acc[3][k])) * FMA(-tp->h1[n][k], tp->h1[n][k], (REAL)1);
        for (int c4 = 0; c4 < 4; ++c4) for (int i = 0; i < n_in; ++i) acc[c4][i] = 0;
        for (int j = 0; j < W; ++j) {
            const float* w = m->W1[n] + (size_t)j * n_in;
            REAL* a = acc[j & 3];
            const REAL dj = d1[j];
            for (int i = 0; i < n_in; ++i) a[i] = FMA((REAL)w[i], dj, a[i]);
        }
        for (int i = 0; i < n_in; ++i) lzn[n][i] = (acc[0][i] + acc[1][i]) + (acc[2][i] + acc[3][i]);
    }
    for (int i = 0; i < n_in; ++i) lz[i] = lzn[0][i] + lzn[1][i];
}

/* ---- rigid body helpers ------------------------------------------------------ */
static void FN(rotmat)(const REAL q[4], REAL R[3][3]) {
    const REAL w = q[0], x = q[1], y = q[2], z = q[3];
    const REAL xx = x * x, yy = y * y, zz = z * z, xy = x * y, xz = x * z, yz = y * z, wx = w * x, wy = w * y, wz = w * z;
    R[0][0] = FMA((REAL)-2, yy + zz, (REAL)1); R[0][1] = (REAL)2 * (xy - wz); R[0][2] = (REAL)2 * (xz + wy);
    R[1][0] = (REAL)2 * (xy + wz); R[1][1] = FMA((REAL)-2, xx + zz, (REAL)1); R[1][2] = (REAL)2 * (yz - wx);
    R[2][0] = (REAL)2 * (xz - wy); R[2][1] = (REAL)2 * (yz + wx); R[2][2] = FMA((REAL)-2, xx + yy, (REAL)1);
}

/* quaternion error vector part of conj(rq) (x) qq */
static void FN(quat_err)(const REAL* rq, const REAL* qq, REAL e[3]) {
    e[0] = FMA(rq[3], qq[2], FMA(-rq[2], qq[3], FMA(-rq[1], qq[0], rq[0] * qq[1])));
    e[1] = FMA(-rq[3], qq[1], FMA(-rq[2], qq[0], FMA(rq[1], qq[3], rq[0] * qq[2])));
    e[2] = FMA(-rq[3], qq[0], FMA(rq[2], qq[1], FMA(-rq[1], qq[2], rq[0] * qq[3])));
}

/* Per-step tape for the adjoint */
typedef struct {
    REAL x[NX];      /* state entering the step             */
    REAL xn[NX];     /* state leaving the step (normalised) */
    REAL rn;         /* 1/|q~|                               */
    REAL sig[6];     /* diffusion                            */
    REAL dsg[6];     /* sig0 * sigmoid(s): d sig / d s       */
    REAL disc;       /* gamma^t                              */
    OTAPE mlp;
} CAT(ostep_, SUFFIX);
#define OSTEP CAT(ostep_, SUFFIX)

/* One Euler-Maruyama step + its stage cost ([SPEC] "Drift", "Diffusion", "EM step", "Cost").
 * Returns the undiscounted stage cost. */
static REAL FN(step_fwd)(const sdempc_config* c, const OMODEL* m, OSTEP* st, const REAL* u,
                         const REAL* uprev, const REAL* xr, const REAL* xi, REAL dt, REAL sdt) {
    const REAL* x = st->x;
    const REAL *v = x + 3, *q = x + 6, *w = x + 10;
    const int nu = m->nu;
    REAL R[3][3];
    FN(rotmat)(q, R);
    OTAPE* tp = &st->mlp;
    for (int i = 0; i < 3; ++i) tp->z[i] = FMA(R[2][i], v[2], FMA(R[1][i], v[1], R[0][i] * v[0]));
    for (int i = 0; i < 3; ++i) tp->z[3 + i] = w[i];
    for (int i = 0; i < nu; ++i) tp->z[6 + i] = u[i];
    FN(mlp_fwd)(m, tp);
    const REAL* r = tp->out[0];
    REAL Tsum = 0, Mb[3];
    for (int k = 0; k < 3; ++k) Mb[k] = m->J[k] * r[3 + k];
    for (int i = 0; i < nu; ++i) {
        REAL T = m->kT * (u[i] * u[i]);
        Tsum = (i == 0) ? T : Tsum + T;
        for (int k = 0; k < 3; ++k) Mb[k] = FMA(m->mixer[k][i], T, Mb[k]);
    }
    const REAL fb0 = r[0], fb1 = r[1], fb2 = FMA(-Tsum, m->inv_m, r[2]);
    REAL acc[3];
    for (int i = 0; i < 3; ++i) acc[i] = FMA(R[i][2], fb2, FMA(R[i][1], fb1, R[i][0] * fb0));
    acc[2] = acc[2] + m->gravity;
    const REAL qd0 = (REAL)-0.5 * FMA(q[3], w[2], FMA(q[2], w[1], q[1] * w[0]));
    const REAL qd1 = (REAL)0.5 * FMA(-q[3], w[1], FMA(q[2], w[2], q[0] * w[0]));
    const REAL qd2 = (REAL)0.5 * FMA(q[3], w[0], FMA(-q[1], w[2], q[0] * w[1]));
    const REAL qd3 = (REAL)0.5 * FMA(-q[2], w[0], FMA(q[1], w[1], q[0] * w[2]));
    const REAL gy0 = (m->Jd[0] * w[1]) * w[2], gy1 = (m->Jd[1] * w[2]) * w[0], gy2 = (m->Jd[2] * w[0]) * w[1];
    const REAL wd[3] = {m->Jinv[0] * (Mb[0] - gy0), m->Jinv[1] * (Mb[1] - gy1), m->Jinv[2] * (Mb[2] - gy2)};
    REAL sig2 = 0;
    for (int i = 0; i < 6; ++i) {
        st->sig[i] = m->sig0[i] * M_SOFTPLUS(tp->out[1][i]);
        st->dsg[i] = m->sig0[i] * M_SIGMOID(tp->out[1][i]);
        sig2 = (i == 0) ? st->sig[0] * st->sig[0] : FMA(st->sig[i], st->sig[i], sig2);
    }
    REAL* xn = st->xn;
    for (int i = 0; i < 3; ++i) xn[i] = FMA(v[i], dt, x[i]);
    for (int i = 0; i < 3; ++i) xn[3 + i] = FMA(st->sig[i] * xi[i], sdt, FMA(acc[i], dt, v[i]));
    const REAL qt0 = FMA(qd0, dt, q[0]), qt1 = FMA(qd1, dt, q[1]), qt2 = FMA(qd2, dt, q[2]), qt3 = FMA(qd3, dt, q[3]);
    const REAL n2 = FMA(qt3, qt3, FMA(qt2, qt2, FMA(qt1, qt1, qt0 * qt0)));
    st->rn = M_RSQRT1(n2);
    xn[6] = qt0 * st->rn; xn[7] = qt1 * st->rn; xn[8] = qt2 * st->rn; xn[9] = qt3 * st->rn;
    for (int i = 0; i < 3; ++i) xn[10 + i] = FMA(st->sig[3 + i] * xi[3 + i], sdt, FMA(wd[i], dt, w[i]));
    /* stage cost on x_{t+1}, u_t */
    REAL l = 0;
    for (int i = 0; i < 3; ++i) { REAL e = xn[i] - xr[i]; l = FMA((REAL)c->perr[i] * e, e, l); }
    for (int i = 0; i < 3; ++i) { REAL e = xn[3 + i] - xr[3 + i]; l = FMA((REAL)c->verr[i] * e, e, l); }
    REAL eq[3];
    FN(quat_err)(xr + 6, xn + 6, eq);
    for (int i = 0; i < 3; ++i) l = FMA((REAL)c->qerr[i] * eq[i], eq[i], l);
    for (int i = 0; i < 3; ++i) { REAL e = xn[10 + i] - xr[10 + i]; l = FMA((REAL)c->werr[i] * e, e, l); }
    for (int i = 0; i < nu; ++i) {
        REAL du = u[i] - (REAL)c->uref[i], ds = u[i] - uprev[i];
        l = FMA((REAL)c->uerr * du, du, l);
        l = FMA((REAL)c->u_slew_coeff * ds, ds, l);
    }
    l = FMA((REAL)c->res_mult, sig2, l);
    return l;
}

/* Adjoint of step_fwd.  lam holds dJ/dx_{t+1} on entry (WITHOUT this stage's
 * direct cost term) and dJ/dx_t on exit; gu[nu] receives dJ/du_t of this stage
 * (dynamics + uerr + this stage's slew term wrt u_t); gprev[nu] the slew term
 * wrt u_{t-1}. */
static void FN(step_bwd)(const sdempc_config* c, const OMODEL* m, const OSTEP* st, const REAL* u,
                         const REAL* uprev, const REAL* xr, const REAL* xi, REAL dt, REAL sdt,
                         REAL lam[NX], REAL* gu, REAL* gprev) {
    const REAL* x = st->x;
    const REAL *v = x + 3, *q = x + 6, *w = x + 10;
    const REAL* xn = st->xn;
    const int nu = m->nu;
    const REAL g2 = (REAL)2 * st->disc;
    /* direct cost on x_{t+1} */
    for (int i = 0; i < 3; ++i) {
        lam[i] = FMA(g2 * (REAL)c->perr[i], xn[i] - xr[i], lam[i]);
        lam[3 + i] = FMA(g2 * (REAL)c->verr[i], xn[3 + i] - xr[3 + i], lam[3 + i]);
        lam[10 + i] = FMA(g2 * (REAL)c->werr[i], xn[10 + i] - xr[10 + i], lam[10 + i]);
    }
    {
        const REAL* rq = xr + 6;
        REAL e[3];
        FN(quat_err)(rq, xn + 6, e);
        const REAL k0 = (g2 * (REAL)c->qerr[0]) * e[0], k1 = (g2 * (REAL)c->qerr[1]) * e[1], k2 = (g2 * (REAL)c->qerr[2]) * e[2];
        lam[6] = lam[6] - FMA(rq[3], k2, FMA(rq[2], k1, rq[1] * k0));
        lam[7] = lam[7] + FMA(rq[2], k2, FMA(-rq[3], k1, rq[0] * k0));
        lam[8] = lam[8] + FMA(-rq[1], k2, FMA(rq[0], k1, rq[3] * k0));
        lam[9] = lam[9] + FMA(rq[0], k2, FMA(rq[1], k1, (-rq[2]) * k0));
    }
    /* normalisation q+ = q~ * rn */
    REAL lqt[4];
    {
        const REAL dot = FMA(xn[9], lam[9], FMA(xn[8], lam[8], FMA(xn[7], lam[7], xn[6] * lam[6])));
        for (int i = 0; i < 4; ++i) lqt[i] = FMA(-xn[6 + i], dot, lam[6 + i]) * st->rn;
    }
    REAL lp[3], lv[3], lw[3], lq[4], la[3], lwd[3], lsig[6];
    const REAL rs2 = g2 * (REAL)c->res_mult;
    for (int i = 0; i < 3; ++i) {
        lp[i] = lam[i];
        lv[i] = FMA(dt, lam[i], lam[3 + i]);
        la[i] = dt * lam[3 + i];
        lw[i] = lam[10 + i];
        lwd[i] = dt * lam[10 + i];
        lsig[i] = FMA(rs2, st->sig[i], (lam[3 + i] * xi[i]) * sdt);
        lsig[3 + i] = FMA(rs2, st->sig[3 + i], (lam[10 + i] * xi[3 + i]) * sdt);
    }
    /* qdot */
    {
        const REAL l0 = dt * lqt[0], l1 = dt * lqt[1], l2 = dt * lqt[2], l3 = dt * lqt[3];
        lq[0] = FMA((REAL)0.5, FMA(w[2], l3, FMA(w[1], l2, w[0] * l1)), lqt[0]);
        lq[1] = FMA((REAL)0.5, FMA(w[1], l3, FMA(-w[2], l2, (-w[0]) * l0)), lqt[1]);
        lq[2] = FMA((REAL)0.5, FMA(-w[0], l3, FMA(w[2], l1, (-w[1]) * l0)), lqt[2]);
        lq[3] = FMA((REAL)0.5, FMA(w[0], l2, FMA(-w[1], l1, (-w[2]) * l0)), lqt[3]);
        lw[0] = FMA((REAL)0.5, FMA(-q[2], l3, FMA(q[3], l2, FMA(q[0], l1, (-q[1]) * l0))), lw[0]);
        lw[1] = FMA((REAL)0.5, FMA(q[1], l3, FMA(q[0], l2, FMA(-q[3], l1, (-q[2]) * l0))), lw[1]);
        lw[2] = FMA((REAL)0.5, FMA(q[0], l3, FMA(-q[1], l2, FMA(q[2], l1, (-q[3]) * l0))), lw[2]);
    }
    /* wdot = Jinv (Mb - gyro) */
    const REAL lMb[3] = {m->Jinv[0] * lwd[0], m->Jinv[1] * lwd[1], m->Jinv[2] * lwd[2]};
    lw[0] = lw[0] - FMA(m->Jd[2] * w[1], lMb[2], (m->Jd[1] * w[2]) * lMb[1]);
    lw[1] = lw[1] - FMA(m->Jd[2] * w[0], lMb[2], (m->Jd[0] * w[2]) * lMb[0]);
    lw[2] = lw[2] - FMA(m->Jd[1] * w[0], lMb[1], (m->Jd[0] * w[1]) * lMb[0]);
    REAL R[3][3];
    FN(rotmat)(q, R);
    const REAL* r = st->mlp.out[0];
    REAL Tsum = 0;
    for (int i = 0; i < nu; ++i) { REAL T = m->kT * (u[i] * u[i]); Tsum = (i == 0) ? T : Tsum + T; }
    const REAL fb[3] = {r[0], r[1], FMA(-Tsum, m->inv_m, r[2])};
    REAL lfb[3];
    for (int j = 0; j < 3; ++j) lfb[j] = FMA(R[2][j], la[2], FMA(R[1][j], la[1], R[0][j] * la[0]));
    REAL lout[2][6];
    for (int k = 0; k < 3; ++k) { lout[0][k] = lfb[k]; lout[0][3 + k] = m->J[k] * lMb[k]; }
    for (int i = 0; i < 6; ++i) lout[1][i] = lsig[i] * st->dsg[i];
    const REAL lTsum = -(lfb[2] * m->inv_m);
    for (int i = 0; i < nu; ++i) {
        const REAL lT = FMA(m->mixer[2][i], lMb[2], FMA(m->mixer[1][i], lMb[1], FMA(m->mixer[0][i], lMb[0], lTsum)));
        gu[i] = (m->kT2 * u[i]) * lT;
    }
    REAL lz[6 + SDEMPC_MAX_NU];
    FN(mlp_bwd)(m, &st->mlp, lout, lz);
    for (int i = 0; i < 3; ++i) lw[i] = lw[i] + lz[3 + i];
    for (int i = 0; i < nu; ++i) gu[i] = gu[i] + lz[6 + i];
    /* vb = R^T v ; acc = R fb : d/dq of sum_ij M_ij R_ij(q), M = la fb^T + v lz^T */
    {
        REAL M[3][3];
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) M[i][j] = FMA(v[i], lz[j], la[i] * fb[j]);
        const REAL qw = q[0], qx = q[1], qy = q[2], qz = q[3];
        const REAL m2x = (REAL)-2 * qx, m2y = (REAL)-2 * qy, m2z = (REAL)-2 * qz;
        const REAL d0 = FMA(qx, M[2][1], FMA(-qy, M[2][0], FMA(-qx, M[1][2], FMA(qz, M[1][0], FMA(qy, M[0][2], (-qz) * M[0][1])))));
        const REAL d1 = FMA(m2x, M[2][2], FMA(qw, M[2][1], FMA(qz, M[2][0], FMA(-qw, M[1][2], FMA(m2x, M[1][1], FMA(qy, M[1][0], FMA(qz, M[0][2], qy * M[0][1])))))));
        const REAL d2 = FMA(m2y, M[2][2], FMA(qz, M[2][1], FMA(-qw, M[2][0], FMA(qz, M[1][2], FMA(qx, M[1][0], FMA(qw, M[0][2], FMA(qx, M[0][1], m2y * M[0][0])))))));
        const REAL d3 = FMA(qy, M[2][1], FMA(qx, M[2][0], FMA(qy, M[1][2], FMA(m2z, M[1][1], FMA(qw, M[1][0], FMA(qx, M[0][2], FMA(-qw, M[0][1], m2z * M[0][0])))))));
        lq[0] = FMA((REAL)2, d0, lq[0]); lq[1] = FMA((REAL)2, d1, lq[1]);
        lq[2] = FMA((REAL)2, d2, lq[2]); lq[3] = FMA((REAL)2, d3, lq[3]);
    }
    for (int i = 0; i < 3; ++i) lv[i] = FMA(R[i][2], lz[2], FMA(R[i][1], lz[1], FMA(R[i][0], lz[0], lv[i])));
    /* direct u terms */
    for (int i = 0; i < nu; ++i) {
        const REAL ds = (g2 * (REAL)c->u_slew_coeff) * (u[i] - uprev[i]);
        gu[i] = FMA(g2 * (REAL)c->uerr, u[i] - (REAL)c->uref[i], gu[i]) + ds;
        gprev[i] = -ds;
    }
    for (int i = 0; i < 3; ++i) { lam[i] = lp[i]; lam[3 + i] = lv[i]; lam[10 + i] = lw[i]; }
    for (int i = 0; i < 4; ++i) lam[6 + i] = lq[i];
}

/* ---- reference window -------------------------------------------------------- */
/* quaternion renormalisation with IEEE sqrt and division (off the per-step path) */
static void FN(quat_renorm)(REAL* q) {
    const REAL n2 = FMA(q[3], q[3], FMA(q[2], q[2], FMA(q[1], q[1], q[0] * q[0])));
    const REAL inv = (REAL)1 / R_SQRT(n2);
    for (int i = 0; i < 4; ++i) q[i] = q[i] * inv;
}

/* linear interpolation per column, clamped at both ends, quaternion renormalised
 * ([SPEC] "Reference window"; state_from_traj, sde_control.py:206, 694) */
static void FN(traj_interp)(const float* tab, int T, REAL t, REAL* out) {
    int lo = 0, hi = T - 1;
    if (t <= (REAL)tab[0]) {
        for (int i = 0; i < NX; ++i) out[i] = tab[1 + i];
    } else if (t >= (REAL)tab[(size_t)hi * 14]) {
        for (int i = 0; i < NX; ++i) out[i] = tab[(size_t)hi * 14 + 1 + i];
    } else {
        while (hi - lo > 1) { /* invariant: t[lo] <= t < t[hi] */
            int mid = (lo + hi) >> 1;
            if ((REAL)tab[(size_t)mid * 14] <= t) lo = mid; else hi = mid;
        }
        const float *a = tab + (size_t)lo * 14, *b = a + 14;
        const REAL al = (t - (REAL)a[0]) / ((REAL)b[0] - (REAL)a[0]);
        for (int i = 0; i < NX; ++i) out[i] = FMA(al, (REAL)b[1 + i] - (REAL)a[1 + i], (REAL)a[1 + i]);
    }
    FN(quat_renorm)(out + 6);
}

/* ---- soft input-rate constraint (include/sdempc.h) -------------------------------------
 * A function of the control sequence only: J_rate = sum_i w_t e_i^2 over i = (t, k), accumulated as 32
 * strided partials + butterfly; dJ_rate/du_i = 2 w_t e_i - 2 w_{t+1} e_{i+nu}; w_t = coeff * discount^t
 * (running product), e_i = violation of [lo, hi] by u_t[k] - u_{t-1}[k].  Added to every particle's cost
 * and gradient. */
static REAL FN(butterfly)(REAL* p);
static REAL FN(rate_violation)(const sdempc_config* c, int nu, const REAL* u, const REAL* uprev0, int t, int k) {
    const REAL ds = u[t * nu + k] - (t == 0 ? uprev0[k] : u[(t - 1) * nu + k]);
    const REAL hi = (REAL)c->u_slew_hi[k], lo = (REAL)c->u_slew_lo[k];
    return ds > hi ? ds - hi : (ds < lo ? ds - lo : (REAL)0);
}
static REAL FN(rate_terms)(const sdempc_config* c, int nu, const REAL* u, const REAL* uprev0, REAL* rgrad) {
    const int H = c->horizon, n = H * nu;
    REAL w[SDEMPC_MAX_H], part[32];
    REAL disc = 1;
    for (int t = 0; t < H; ++t) { w[t] = disc * (REAL)c->u_slew_constr_coeff; disc = disc * (REAL)c->discount; }
    for (int l = 0; l < 32; ++l) part[l] = 0;
    for (int i = 0; i < n; ++i) {
        const int t = i / nu, k = i % nu;
        const REAL e = FN(rate_violation)(c, nu, u, uprev0, t, k);
        part[i & 31] = FMA(w[t] * e, e, part[i & 31]);
        if (rgrad) {
            REAL a = ((REAL)2 * w[t]) * e;
            if (t + 1 < H) a = a - ((REAL)2 * w[t + 1]) * FN(rate_violation)(c, nu, u, uprev0, t + 1, k);
            rgrad[i] = a;
        }
    }
    return FN(butterfly)(part);
}

/* ---- rollout: J(u) and dJ/du --------------------------------------------------- */
/* xref[H+1][13] internal frame; xi[P][H][6]; grad[H][nu] or NULL; xmean[H+1][13] or NULL. */
static REAL FN(rollout)(const sdempc_config* c, const OMODEL* m, const REAL* x0, const REAL* u,
                        const REAL* uprev0, const REAL* xref, const REAL* xi, REAL* grad,
                        REAL* xmean, OSTEP* steps) {
    const int H = c->horizon, nu = m->nu, P = c->num_particles;
    const REAL invP = (REAL)1 / (REAL)P;
    REAL J = 0;
    REAL gpart[NUH], rgrad[NUH];
    const int rate_on = c->u_slew_constr_coeff != 0.0f;
    const REAL Jrate = rate_on ? FN(rate_terms)(c, nu, u, uprev0, grad ? rgrad : NULL) : (REAL)0;
    if (grad) for (int i = 0; i < H * nu; ++i) grad[i] = 0;
    if (xmean) for (int i = 0; i < (H + 1) * NX; ++i) xmean[i] = 0;
    for (int p = 0; p < P; ++p) {
        REAL Jp = 0, disc = 1;
        for (int i = 0; i < NX; ++i) steps[0].x[i] = x0[i];
        for (int t = 0; t < H; ++t) {
            OSTEP* st = &steps[t];
            st->disc = disc;
            const REAL* up = (t == 0) ? uprev0 : u + (size_t)(t - 1) * nu;
            const REAL dt = (REAL)c->dt[t], sdt = R_SQRT((REAL)c->dt[t]);
            const REAL l = FN(step_fwd)(c, m, st, u + (size_t)t * nu, up, xref + (size_t)(t + 1) * NX,
                                        xi + ((size_t)p * H + t) * 6, dt, sdt);
            Jp = FMA(disc, l, Jp);
            if (t + 1 < H) for (int i = 0; i < NX; ++i) steps[t + 1].x[i] = st->xn[i];
            disc = disc * (REAL)c->discount;
        }
        if (rate_on) Jp = Jp + Jrate;
        J = (p == 0) ? Jp : J + Jp;
        if (xmean) {
            for (int i = 0; i < NX; ++i) xmean[i] = (p == 0) ? x0[i] : xmean[i] + x0[i];
            for (int t = 0; t < H; ++t)
                for (int i = 0; i < NX; ++i) {
                    REAL* d = &xmean[(size_t)(t + 1) * NX + i];
                    *d = (p == 0) ? steps[t].xn[i] : *d + steps[t].xn[i];
                }
        }
        if (grad) {
            REAL lam[NX], gu[SDEMPC_MAX_NU], gp[SDEMPC_MAX_NU], gn[SDEMPC_MAX_NU];
            for (int i = 0; i < NX; ++i) lam[i] = 0;
            for (int i = 0; i < nu; ++i) gp[i] = 0;
            for (int t = H - 1; t >= 0; --t) {
                const REAL* up = (t == 0) ? uprev0 : u + (size_t)(t - 1) * nu;
                const REAL dt = (REAL)c->dt[t], sdt = R_SQRT((REAL)c->dt[t]);
                FN(step_bwd)(c, m, &steps[t], u + (size_t)t * nu, up, xref + (size_t)(t + 1) * NX,
                             xi + ((size_t)p * H + t) * 6, dt, sdt, lam, gu, gn);
                for (int i = 0; i < nu; ++i) { gpart[t * nu + i] = gu[i] + gp[i]; gp[i] = gn[i]; }
            }
            if (rate_on) for (int i = 0; i < H * nu; ++i) gpart[i] = gpart[i] + rgrad[i];
            for (int i = 0; i < H * nu; ++i) grad[i] = (p == 0) ? gpart[i] : grad[i] + gpart[i];
        }
    }
    if (grad) for (int i = 0; i < H * nu; ++i) grad[i] = grad[i] * invP;
    if (xmean) {
        for (int t = 0; t <= H; ++t) {
            REAL* r = xmean + (size_t)t * NX;
            for (int i = 0; i < NX; ++i) r[i] = r[i] * invP;
            FN(quat_renorm)(r + 6);
        }
    }
    return J * invP;
}

/* ---- per-problem front end ------------------------------------------------------ */
typedef struct {
    sdempc_config cfg;
    OMODEL model;
    void* blob;
    float* traj_ext; /* [T][14] external frame */
    float* traj_int; /* [T][14] internal frame */
    int T;
} CAT(ohandle_, SUFFIX);
#define OHANDLE CAT(ohandle_, SUFFIX)

static REAL FN(clip)(REAL v, REAL lo, REAL hi) { v = v < lo ? lo : v; return v > hi ? hi : v; }

/* SPEC-ARITH warp-shaped sum: 32 strided partials, then an xor-butterfly (16,8,4,2,1) */
static REAL FN(butterfly)(REAL* p) {
    for (int off = 16; off >= 1; off >>= 1) {
        REAL q[32];
        for (int l = 0; l < 32; ++l) q[l] = p[l] + p[l ^ off];
        for (int l = 0; l < 32; ++l) p[l] = q[l];
    }
    return p[0];
}

/* Build the internal-frame problem data: x0, reference window, noise. */
static int FN(prepare)(const OHANDLE* h, const REAL* x_ext, const REAL* curr_t, const REAL* xdes,
                       const REAL* xref_win, const uint64_t* rng, const REAL* xi_override,
                       REAL* x0, REAL* xref, REAL* xi) {
    const sdempc_config* c = &h->cfg;
    const int H = c->horizon, P = c->num_particles;
    const int enu = (c->flags & SDEMPC_F_FRAME_ENU) != 0;
    if (enu) FN(enu_ned)(x_ext, x0); else for (int i = 0; i < NX; ++i) x0[i] = x_ext[i];
    if (xref_win) {
        for (int t = 0; t <= H; ++t) {
            if (enu) FN(enu_ned)(xref_win + (size_t)t * NX, xref + (size_t)t * NX);
            else for (int i = 0; i < NX; ++i) xref[(size_t)t * NX + i] = xref_win[(size_t)t * NX + i];
        }
    } else if (curr_t) {
        if (!h->traj_int) return SDEMPC_ESTATE;
        REAL tt = *curr_t;
        for (int t = 0; t <= H; ++t) {
            FN(traj_interp)(h->traj_int, h->T, tt, xref + (size_t)t * NX);
            if (t < H) tt = tt + (REAL)c->dt[t];
        }
    } else if (xdes) {
        REAL xd[NX];
        if (enu) FN(enu_ned)(xdes, xd); else for (int i = 0; i < NX; ++i) xd[i] = xdes[i];
        for (int t = 0; t <= H; ++t) for (int i = 0; i < NX; ++i) xref[(size_t)t * NX + i] = xd[i];
    } else return SDEMPC_EINVAL;
    if (xi_override) for (int i = 0; i < P * H * 6; ++i) xi[i] = xi_override[i];
    else {
        if (!rng) return SDEMPC_EINVAL;
        FN(gen_noise)(rng[0], rng[1], P, H, 0, xi);
    }
    return 0;
}

static void FN(to_ext)(const OHANDLE* h, const REAL* xin, REAL* xout, int n) {
    const int enu = (h->cfg.flags & SDEMPC_F_FRAME_ENU) != 0;
    for (int t = 0; t < n; ++t) {
        if (enu) FN(enu_ned)(xin + (size_t)t * NX, xout + (size_t)t * NX);
        else for (int i = 0; i < NX; ++i) xout[(size_t)t * NX + i] = xin[(size_t)t * NX + i];
    }
}

typedef struct {
    REAL avg_linesearch, stepsize, num_steps, grad_sqr, avg_stepsize, init_cost, opt_cost, solve_time_us;
} CAT(oinfo_, SUFFIX);
#define OINFO CAT(oinfo_, SUFFIX)

/* The APG solve of one problem in the internal frame ([SPEC] "APG").
 * plan[H][nu] in/out, xevol[H+1][13] out (internal frame), trace[max_iter][8] or NULL.
 *
 * TEACHER-FORCED mode (forced != NULL; SURVEY.md section 7 "decision-trace teacher-forced mode"): the discrete
 * decisions of every iteration -- the number of line-search trials and accept / reject -- and the iteration count
 * are taken from another solver's decision trace `forced[forced_iters][8]` instead of being decided here, so the
 * sequence of iterates y_k, x_k is the one that solver followed and every continuous quantity (f_y, J_trial, step
 * size, J_x, |g|^2) can be compared with it at a tolerance although a free-running solve would branch differently
 * after the first flipped decision.  own[it][4] = (the trial count this oracle would have chosen given the same
 * trials, its own accept decision, Armijo margin of the last forced trial (>= 0: passes), accept margin J_x - J_trial). */
static void FN(apg)(const OHANDLE* h, const REAL* x0, const REAL* xref, const REAL* xi, REAL* plan,
                    REAL* xevol, OINFO* info, REAL* trace, OSTEP* steps, const REAL* forced, int forced_iters, REAL* own) {
    const sdempc_config* c = &h->cfg;
    const OMODEL* m = &h->model;
    const int H = c->horizon, nu = m->nu, n = H * nu;
    REAL xk[NUH], yk[NUH], g[NUH], xp[NUH], uprev[SDEMPC_MAX_NU], part[32];
    for (int i = 0; i < nu; ++i) uprev[i] = plan[i];
    for (int t = 0; t < H; ++t) {
        int ts = (c->flags & SDEMPC_F_NO_SHIFT) ? t : (t + 1 < H ? t + 1 : H - 1);
        for (int i = 0; i < nu; ++i)
            xk[t * nu + i] = FN(clip)(plan[ts * nu + i], (REAL)c->u_lo[i], (REAL)c->u_hi[i]);
    }
    for (int i = 0; i < n; ++i) yk[i] = xk[i];
    REAL s = info->stepsize > 0 ? info->stepsize : (REAL)c->init_stepsize;
    REAL Jx = 0, Jp = 0, fy = 0, gsq = 0, sum_ls = 0, sum_s = 0, init_cost = 0;
    int k = 1, no_improve = 0, it = 0;
    for (;;) {
        ++it;
        fy = FN(rollout)(c, m, x0, yk, uprev, xref, xi, g, NULL, steps);
        if (it == 1) { Jx = fy; init_cost = fy; }
        for (int l = 0; l < 32; ++l) part[l] = 0;
        for (int i = 0; i < n; ++i) part[i & 31] = FMA(g[i], g[i], part[i & 31]);
        gsq = FN(butterfly)(part);
        if (c->reset_option == 1) { s = s * (REAL)c->increase_factor; s = s > (REAL)c->max_stepsize ? (REAL)c->max_stepsize : s; }
        int ok = 0, n_ls = 0;
        const REAL* fr = forced ? forced + (size_t)(it - 1) * SDEMPC_TRACE_W : NULL;
        const int n_forced = fr ? (int)fr[3] : 0;
        int own_nls = 0;
        REAL margin_ls = 0;
        for (int j = 0; j <= c->maxls; ++j) {
            for (int l = 0; l < 32; ++l) part[l] = 0;
            for (int i = 0; i < n; ++i) {
                xp[i] = FN(clip)(FMA(-s, g[i], yk[i]), (REAL)c->u_lo[i % nu], (REAL)c->u_hi[i % nu]);
                part[i & 31] = FMA(g[i], xp[i] - yk[i], part[i & 31]);
            }
            const REAL dec = FN(butterfly)(part);
            Jp = FN(rollout)(c, m, x0, xp, uprev, xref, xi, NULL, NULL, steps);
            n_ls = j + 1;
            ok = (Jp <= FMA((REAL)c->coef, dec, fy));
            margin_ls = FMA((REAL)c->coef, dec, fy) - Jp;
            if (fr) {   /* teacher forced: evaluate exactly the forced trials, remember where this oracle would have stopped */
                if (ok && own_nls == 0) own_nls = j + 1;
                if (j + 1 >= n_forced) break;
                s = s * (REAL)c->decrease_factor;
                continue;
            }
            if (ok) break;
            if (j < c->maxls) s = s * (REAL)c->decrease_factor;
        }
        sum_ls = sum_ls + (REAL)n_ls; sum_s = sum_s + s;
        int accept = ok && (Jp <= Jx), converged = 0;
        if (fr) {
            if (own) {
                REAL* o = own + (size_t)(it - 1) * 4;
                o[0] = (REAL)(own_nls ? own_nls : (n_forced <= c->maxls ? n_forced + 1 : n_forced));
                o[1] = (REAL)accept; o[2] = margin_ls; o[3] = Jx - Jp;
            }
            accept = fr[4] != 0;
        }
        if (accept) {
            REAL beta = (REAL)k / (REAL)(k + 3);
            if (c->moment_scale != 0.0f) {   /* [SPEC] adaptive momentum: beta_init scaled by 1 / moment_scale per accepted step, capped at 1 */
                beta = (REAL)c->beta_init;
                for (int i = 1; i < k && beta < (REAL)1; ++i) { beta = beta / (REAL)c->moment_scale; beta = beta > (REAL)1 ? (REAL)1 : beta; }
            }
            for (int i = 0; i < n; ++i) {
                yk[i] = FN(clip)(FMA(beta, xp[i] - xk[i], xp[i]), (REAL)c->u_lo[i % nu], (REAL)c->u_hi[i % nu]);
                xk[i] = xp[i];
            }
            const REAL Jprev = Jx;
            Jx = Jp; ++k; no_improve = 0;
            const REAL tol = (REAL)c->atol + (REAL)c->rtol * R_FABS(Jprev);
            converged = (R_FABS(Jprev - Jx) <= tol) || (Jx <= (REAL)c->atol);
        } else {
            for (int i = 0; i < n; ++i) yk[i] = xk[i];
            k = 1; ++no_improve;
        }
        if (trace) {
            REAL* tr = trace + (size_t)(it - 1) * SDEMPC_TRACE_W;
            tr[0] = fy; tr[1] = Jp; tr[2] = s; tr[3] = (REAL)n_ls; tr[4] = (REAL)accept; tr[5] = Jx; tr[6] = gsq; tr[7] = (REAL)k;
        }
        if (fr) { if (it >= forced_iters) break; else continue; }
        if (it >= c->max_iter || no_improve >= c->max_no_improvement_iter || converged || !(fy == fy)) break;
    }
    FN(rollout)(c, m, x0, xk, uprev, xref, xi, NULL, xevol, steps);
    for (int i = 0; i < n; ++i) plan[i] = xk[i];
    info->avg_linesearch = sum_ls / (REAL)it; info->stepsize = s; info->num_steps = (REAL)it;
    info->grad_sqr = gsq; info->avg_stepsize = sum_s / (REAL)it; info->init_cost = init_cost;
    info->opt_cost = (Jx == Jx) ? Jx : (REAL)INFINITY;
}

/* ---- exported API (array element type = REAL) ------------------------------------ */
int FN(create)(const sdempc_config* cfg, const void* blob, size_t nbytes, void** out) {
    OHANDLE* h = (OHANDLE*)calloc(1, sizeof(OHANDLE));
    if (!h) return SDEMPC_ENOMEM;
    h->cfg = *cfg;
    h->blob = malloc(nbytes);
    memcpy(h->blob, blob, nbytes);
    if (FN(parse_model)(h->blob, nbytes, &h->model) || cfg->nu != h->model.nu || cfg->horizon < 1 ||
        cfg->horizon > SDEMPC_MAX_H || cfg->num_particles < 1) {
        free(h->blob); free(h);
        return SDEMPC_EINVAL;
    }
    *out = h;
    return 0;
}

void FN(destroy)(void* hv) {
    OHANDLE* h = (OHANDLE*)hv;
    if (!h) return;
    for (int n = 0; n < 2; ++n) { free(h->model.W1T[n]); free(h->model.W2T[n]); }
    free(h->blob); free(h->traj_ext); free(h->traj_int); free(h);
}

int FN(set_trajectory)(void* hv, const float* table, int T) {
    OHANDLE* h = (OHANDLE*)hv;
    if (T < 2) return SDEMPC_EINVAL;
    free(h->traj_ext); free(h->traj_int);
    h->traj_ext = (float*)malloc(sizeof(float) * 14 * (size_t)T);
    h->traj_int = (float*)malloc(sizeof(float) * 14 * (size_t)T);
    memcpy(h->traj_ext, table, sizeof(float) * 14 * (size_t)T);
    const int enu = (h->cfg.flags & SDEMPC_F_FRAME_ENU) != 0;
    for (int r = 0; r < T; ++r) {
        const float* a = table + (size_t)r * 14;
        float* b = h->traj_int + (size_t)r * 14;
        b[0] = a[0];
        if (r > 0 && !(a[0] > a[-14])) return SDEMPC_EINVAL;
        if (enu) {
            /* table rows are converted in float32 by the same enu_ned statement on both
             * the f32 and f64 oracles and in the CUDA library's host code */
            const float s = 0.70710678118654752440f;
            float c0 = s * (a[7] + a[10]), c1 = s * (a[8] + a[9]), c2 = s * (a[8] - a[9]), c3 = s * (a[7] - a[10]);
            if (c0 < 0) { c0 = -c0; c1 = -c1; c2 = -c2; c3 = -c3; }
            b[1] = a[2]; b[2] = a[1]; b[3] = -a[3];
            b[4] = a[5]; b[5] = a[4]; b[6] = -a[6];
            b[7] = c0; b[8] = c1; b[9] = c2; b[10] = c3;
            b[11] = a[11]; b[12] = -a[12]; b[13] = -a[13];
        } else {
            for (int i = 0; i < NX; ++i) b[1 + i] = a[1 + i];
        }
        /* keep consecutive quaternions in the same hemisphere so that linear
         * interpolation + renormalisation is well defined */
        if (r > 0) {
            float d = 0;
            for (int i = 7; i < 11; ++i) d += b[i] * b[i - 14];
            if (d < 0) for (int i = 7; i < 11; ++i) b[i] = -b[i];
        }
    }
    h->T = T;
    return 0;
}

/* external-frame interpolation (what the node calls state_from_traj) */
int FN(state_from_traj)(void* hv, const REAL* t, int n, REAL* out) {
    OHANDLE* h = (OHANDLE*)hv;
    if (!h->traj_ext) return SDEMPC_ESTATE;
    for (int i = 0; i < n; ++i) FN(traj_interp)(h->traj_ext, h->T, t[i], out + (size_t)i * NX);
    return 0;
}

/* internal-frame trajectory table (for tests) */
const float* FN(traj_internal)(void* hv) { return ((OHANDLE*)hv)->traj_int; }

void FN(enu2ned)(const REAL* x, REAL* o, int n) {
    for (int i = 0; i < n; ++i) FN(enu_ned)(x + (size_t)i * NX, o + (size_t)i * NX);
}

void FN(noise)(uint64_t seed, uint64_t tick, int P, int H, int sub0, REAL* xi) {
    FN(gen_noise)(seed, tick, P, H, (uint32_t)sub0, xi);
}

/* elementary functions, exported for accuracy tests: which = 0 tanh, 1 softplus, 2 sigmoid, 3 rsqrt_near1 */
void FN(elementary)(int which, const REAL* x, REAL* y, int n) {
    for (int i = 0; i < n; ++i)
        y[i] = which == 0 ? M_TANH(x[i]) : which == 1 ? M_SOFTPLUS(x[i]) : which == 2 ? M_SIGMOID(x[i]) : M_RSQRT1(x[i]);
}

int FN(reset)(void* hv, int B, REAL* u_plan, REAL* info /*[B][8]*/) {
    OHANDLE* h = (OHANDLE*)hv;
    const sdempc_config* c = &h->cfg;
    for (int b = 0; b < B; ++b) {
        for (int t = 0; t < c->horizon; ++t)
            for (int i = 0; i < c->nu; ++i)
                u_plan[((size_t)b * c->horizon + t) * c->nu + i] = FN(clip)((REAL)c->uref[i], (REAL)c->u_lo[i], (REAL)c->u_hi[i]);
        for (int i = 0; i < 8; ++i) info[(size_t)b * 8 + i] = 0;
        info[(size_t)b * 8 + 1] = (REAL)c->init_stepsize;
    }
    return 0;
}

/* value_and_grad at u for B problems (mirrors sdempc_rollout) */
int FN(rollout_batch)(void* hv, int B, const REAL* x, const REAL* curr_t, const REAL* xdes, const REAL* xref_win,
                      const uint64_t* rng, const REAL* xi_override, const REAL* u, const REAL* u_prev,
                      REAL* cost, REAL* grad, REAL* x_evol) {
    OHANDLE* h = (OHANDLE*)hv;
    const sdempc_config* c = &h->cfg;
    const int H = c->horizon, nu = c->nu, P = c->num_particles;
    int rc = 0;
#pragma omp parallel for schedule(dynamic, 1)
    for (int b = 0; b < B; ++b) {
        OSTEP* steps = (OSTEP*)malloc(sizeof(OSTEP) * (size_t)H);
        REAL* xi = (REAL*)malloc(sizeof(REAL) * (size_t)P * H * 6);
        REAL x0[NX], xref[(SDEMPC_MAX_H + 1) * NX], xm[(SDEMPC_MAX_H + 1) * NX];
        int e = FN(prepare)(h, x + (size_t)b * NX, curr_t ? curr_t + b : NULL, xdes ? xdes + (size_t)b * NX : NULL,
                            xref_win ? xref_win + (size_t)b * (H + 1) * NX : NULL, rng ? rng + 2 * (size_t)b : NULL,
                            xi_override ? xi_override + (size_t)b * P * H * 6 : NULL, x0, xref, xi);
        if (e) { rc = e; }
        else {
            cost[b] = FN(rollout)(c, &h->model, x0, u + (size_t)b * H * nu, u_prev + (size_t)b * nu, xref, xi,
                                  grad ? grad + (size_t)b * H * nu : NULL, x_evol ? xm : NULL, steps);
            if (x_evol) FN(to_ext)(h, xm, x_evol + (size_t)b * (H + 1) * NX, H + 1);
        }
        free(steps); free(xi);
    }
    return rc;
}

/* batched solve (mirrors sdempc_solve_ex); info[B][8], trace[B][max_iter][8] or NULL.
 * forced[B][max_iter][8] + forced_iters[B] (both or neither): teacher-forced replay of another solver's decision
 * trace (see apg); own[B][max_iter][4] receives this oracle's own decisions and margins. */
int FN(solve_forced)(void* hv, int B, const REAL* x, const REAL* curr_t, const REAL* xdes, const REAL* xref_win,
                     const uint64_t* rng, REAL* u_plan, REAL* x_evol, REAL* info, const REAL* xi_override, REAL* trace,
                     const REAL* forced, const REAL* forced_iters, REAL* own);
int FN(solve)(void* hv, int B, const REAL* x, const REAL* curr_t, const REAL* xdes, const REAL* xref_win,
              const uint64_t* rng, REAL* u_plan, REAL* x_evol, REAL* info, const REAL* xi_override, REAL* trace) {
    return FN(solve_forced)(hv, B, x, curr_t, xdes, xref_win, rng, u_plan, x_evol, info, xi_override, trace, NULL, NULL, NULL);
}
int FN(solve_forced)(void* hv, int B, const REAL* x, const REAL* curr_t, const REAL* xdes, const REAL* xref_win,
                     const uint64_t* rng, REAL* u_plan, REAL* x_evol, REAL* info, const REAL* xi_override, REAL* trace,
                     const REAL* forced, const REAL* forced_iters, REAL* own) {
    OHANDLE* h = (OHANDLE*)hv;
    const sdempc_config* c = &h->cfg;
    const int H = c->horizon, nu = c->nu, P = c->num_particles;
    int rc = 0;
#pragma omp parallel for schedule(dynamic, 1)
    for (int b = 0; b < B; ++b) {
        OSTEP* steps = (OSTEP*)malloc(sizeof(OSTEP) * (size_t)H);
        REAL* xi = (REAL*)malloc(sizeof(REAL) * (size_t)P * H * 6);
        REAL x0[NX], xref[(SDEMPC_MAX_H + 1) * NX], xm[(SDEMPC_MAX_H + 1) * NX];
        int e = FN(prepare)(h, x + (size_t)b * NX, curr_t ? curr_t + b : NULL, xdes ? xdes + (size_t)b * NX : NULL,
                            xref_win ? xref_win + (size_t)b * (H + 1) * NX : NULL, rng ? rng + 2 * (size_t)b : NULL,
                            xi_override ? xi_override + (size_t)b * P * H * 6 : NULL, x0, xref, xi);
        if (e) { rc = e; }
        else {
            OINFO inf;
            memset(&inf, 0, sizeof(inf));
            inf.stepsize = info[(size_t)b * 8 + 1];
            const int fi = forced_iters ? (int)forced_iters[b] : 0;
            FN(apg)(h, x0, xref, xi, u_plan + (size_t)b * H * nu, xm, &inf,
                    trace ? trace + (size_t)b * c->max_iter * SDEMPC_TRACE_W : NULL, steps,
                    (forced && fi > 0) ? forced + (size_t)b * c->max_iter * SDEMPC_TRACE_W : NULL, fi,
                    own ? own + (size_t)b * c->max_iter * 4 : NULL);
            FN(to_ext)(h, xm, x_evol + (size_t)b * (H + 1) * NX, H + 1);
            REAL* o = info + (size_t)b * 8;
            o[0] = inf.avg_linesearch; o[1] = inf.stepsize; o[2] = inf.num_steps; o[3] = inf.grad_sqr;
            o[4] = inf.avg_stepsize; o[5] = inf.init_cost; o[6] = inf.opt_cost; o[7] = 0;
        }
        free(steps); free(xi);
    }
    return rc;
}

/* Monte-Carlo closed loop (mirrors sdempc_closed_loop): plant = same SDE, P = 1,
 * Philox sub-stream 2; one plant EM step of dt[0] per tick. */
int FN(closed_loop)(void* hv, int Rn, int ticks, const REAL* x0s, const REAL* t0s, const uint64_t* rng,
                    REAL* x_hist, REAL* u_hist, REAL* stats) {
    OHANDLE* h = (OHANDLE*)hv;
    const sdempc_config* c = &h->cfg;
    const int H = c->horizon, nu = c->nu, P = c->num_particles;
    if (!h->traj_int) return SDEMPC_ESTATE;
    const int enu = (c->flags & SDEMPC_F_FRAME_ENU) != 0;
#pragma omp parallel for schedule(dynamic, 1)
    for (int r = 0; r < Rn; ++r) {
        OSTEP* steps = (OSTEP*)malloc(sizeof(OSTEP) * (size_t)H);
        REAL* xi = (REAL*)malloc(sizeof(REAL) * (size_t)P * H * 6);
        REAL x[NX], xe[NX], xref[(SDEMPC_MAX_H + 1) * NX], xm[(SDEMPC_MAX_H + 1) * NX];
        REAL plan[NUH];
        OINFO inf;
        memset(&inf, 0, sizeof(inf));
        inf.stepsize = (REAL)c->init_stepsize;
        for (int t = 0; t < H; ++t)
            for (int i = 0; i < nu; ++i) plan[t * nu + i] = FN(clip)((REAL)c->uref[i], (REAL)c->u_lo[i], (REAL)c->u_hi[i]);
        if (enu) FN(enu_ned)(x0s + (size_t)r * NX, x); else for (int i = 0; i < NX; ++i) x[i] = x0s[(size_t)r * NX + i];
        REAL se = 0, me = 0, sc = 0, sn = 0;
        const uint64_t seed = rng[2 * (size_t)r];
        uint64_t tick = rng[2 * (size_t)r + 1];
        const REAL dt0 = (REAL)c->dt[0], sdt0 = R_SQRT((REAL)c->dt[0]);
        for (int k = 0; k < ticks; ++k, ++tick) {
            if (x_hist) { FN(to_ext)(h, x, xe, 1); for (int i = 0; i < NX; ++i) x_hist[((size_t)r * (ticks + 1) + k) * NX + i] = xe[i]; }
            REAL tw = FMA((REAL)k, dt0, t0s[r]);
            for (int t = 0; t <= H; ++t) {
                FN(traj_interp)(h->traj_int, h->T, tw, xref + (size_t)t * NX);
                if (t < H) tw = tw + (REAL)c->dt[t];
            }
            FN(gen_noise)(seed, tick, P, H, 0, xi);
            FN(apg)(h, x, xref, xi, plan, xm, &inf, NULL, steps, NULL, 0, NULL);
            sc = sc + inf.opt_cost; sn = sn + inf.num_steps;
            if (u_hist) for (int i = 0; i < nu; ++i) u_hist[((size_t)r * ticks + k) * nu + i] = plan[i];
            /* plant step */
            REAL xip[6];
            FN(gen_noise)(seed, tick, 1, 1, 2, xip);
            OSTEP ps;
            for (int i = 0; i < NX; ++i) ps.x[i] = x[i];
            ps.disc = 1;
            FN(step_fwd)(c, &h->model, &ps, plan, plan, xref + NX, xip, dt0, sdt0);
            for (int i = 0; i < NX; ++i) x[i] = ps.xn[i];
            /* tracking error against the reference at the new time */
            REAL e2 = 0;
            for (int i = 0; i < 3; ++i) { const REAL d = x[i] - xref[NX + i]; e2 = FMA(d, d, e2); }
            se = se + e2; me = e2 > me ? e2 : me;
        }
        if (x_hist) { FN(to_ext)(h, x, xe, 1); for (int i = 0; i < NX; ++i) x_hist[((size_t)r * (ticks + 1) + ticks) * NX + i] = xe[i]; }
        stats[(size_t)r * 4 + 0] = R_SQRT(se / (REAL)ticks);
        stats[(size_t)r * 4 + 1] = R_SQRT(me);
        stats[(size_t)r * 4 + 2] = sc / (REAL)ticks;
        stats[(size_t)r * 4 + 3] = sn / (REAL)ticks;
        free(steps); free(xi);
    }
    return 0;
}

#undef OMODEL
#undef OTAPE
#undef OSTEP
#undef OHANDLE
#undef OINFO
#undef FMA
#undef R_SQRT
#undef R_FABS
#undef M_TANH
#undef M_SOFTPLUS
#undef M_SIGMOID
#undef M_RSQRT1
#undef M_LOG
#undef M_SINCOS2PI
