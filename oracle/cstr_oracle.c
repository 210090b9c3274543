/*
 * CPU ORACLE, C restatement (TEST INFRASTRUCTURE — the checker and the timed CPU baseline, never
 * the product).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library.
 *
 * Restates, for N independent reactors (reference file:line):
 *   TwoSeriesCSTREnv.step            /root/reference/twoseriescstr.py:394-454
 *   TwoSeriesCSTREnv._dynamics       /root/reference/twoseriescstr.py:456-503
 *   TwoSeriesCSTREnv.compute_reward  /root/reference/twoseriescstr.py:271-392 (the 2 weighted terms)
 *   _normalize/_denormalize          /root/reference/twoseriescstr.py:129-150
 *   generate_initial_state / reset   /root/reference/twoseriescstr.py:167-269 (draws injected as uniforms)
 *   DummyVecEnv.step_wait auto-reset /root/reference/core/common/vec_env/dummy_vec_env.py:56-73
 *
 * Parity status: PINNED.  tests/test_oracle_c.py requires this library (exp_mode=LIBM is compared
 * with tolerance, everything else bit-exactly) to agree with oracle/cstr_oracle.py, which is itself
 * bit-exact against the unmodified reference (tests/test_oracle_vs_reference.py, tests/golden/).
 *
 * Arithmetic: float32 throughout in the f32 entry points, with the reference's constant folding and
 * left-to-right association.  Build with -ffp-contract=off: no mul+add may fuse.
 * The only FMAs are the explicit fmaf() calls inside cstr_expf_shared(), the documented exp
 * algorithm that the CUDA "strict" kernels implement independently (DESIGN.md "shared exp").
 */
#include <math.h>
#include <stdint.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define EXP_LIBM 0   /* libm expf/exp: what a scalar C port of the reference would call            */
#define EXP_SHARED 1 /* cstr_expf_shared: bit-identical to the CUDA strict kernel                  */
#define SQ_POWF 0    /* powf(n, 2): what `np.float32 ** 2` executes in the reference (:291)        */
#define SQ_MUL 1     /* n*n: correctly rounded square, what the CUDA kernels use                    */

/* ---------------------------------------------------------------------------------------------
 * shared exp: exp(x) for x in [-100, 100]
 *   t = x*log2(e); n = RN(t) by the 1.5*2^23 magic add; r = x - n*ln2 (Cody-Waite hi/lo FMAs);
 *   e^r = 1 + (r + r^2*q(r)), q degree 5 (Chebyshev-interpolated on |r| <= 0.35, fp32-rounded);
 *   result = (p * 2^(n>>1)) * 2^(n-(n>>1)).
 * ------------------------------------------------------------------------------------------- */
static inline float bits_to_float(int32_t b) { float f; memcpy(&f, &b, 4); return f; }
static inline int32_t float_to_bits(float f) { int32_t b; memcpy(&b, &f, 4); return b; }

float cstr_expf_shared(float x) {
    const float MAGIC = 12582912.0f; /* 1.5 * 2^23 */
    float t = x * 1.44269504088896341f + MAGIC; /* two roundings: built with -ffp-contract=off */
    int32_t ni = float_to_bits(t) - 0x4B400000;
    float n = t - MAGIC;
    float r = fmaf(n, -0.693145751953125f, x);
    r = fmaf(n, -1.42860682030941723212e-6f, r);
    float q = 0.00019891989359166473f;
    q = fmaf(q, r, 0.001393454847857356f);
    q = fmaf(q, r, 0.008333309553563595f);
    q = fmaf(q, r, 0.04166645556688309f);
    q = fmaf(q, r, 0.1666666716337204f);
    q = fmaf(q, r, 0.5f);
    float rr = r * r;
    float y = fmaf(rr, q, r);
    float p = y + 1.0f;
    int32_t n1 = ni >> 1, n2 = ni - n1;
    float s1 = bits_to_float((n1 + 127) << 23), s2 = bits_to_float((n2 + 127) << 23);
    return (p * s1) * s2;
}

void cstr_expf_shared_array(const float *x, float *y, int64_t n) {
    for (int64_t i = 0; i < n; ++i) y[i] = cstr_expf_shared(x[i]);
}

/* libm powf through a volatile pointer so the compiler cannot rewrite powf(x,2) as x*x */
static float (*volatile powf_ptr)(float, float) = powf;
void cstr_powf2_array(const float *x, float *y, int64_t n) {
    for (int64_t i = 0; i < n; ++i) y[i] = powf_ptr(x[i], 2.0f);
}

static inline float clipf(float x, float lo, float hi) {
    /* np.clip == minimum(maximum(x, lo), hi); NaN propagates */
    if (x != x) return x;
    x = x < lo ? lo : x;
    return x > hi ? hi : x;
}
static inline double clipd(double x, double lo, double hi) {
    if (x != x) return x;
    x = x < lo ? lo : x;
    return x > hi ? hi : x;
}

/* raw bounds (:56-61), float32 */
static const float SLO[4] = {0.0f, 273.15f, 0.0f, 273.15f};
static const float SHI[4] = {0.7f, 400.0f, 0.7f, 400.0f};
#define ALO 30.0f
#define AHI 250.0f
#define MAX_STEPS 400

/* One env, one control interval, float32.  Returns 1 when the NaN path (:415-421) was taken. */
static inline int step_one_f32(float s[4], const float a_in[2], int32_t *step_count, float *reward,
                               int *truncated, float target, int exp_mode, int sq_mode) {
    /* folded constants (App. A) */
    const float QV = (float)(50.0 / 100.0), CF = 0.5f, TF = 320.0f, TCF = 370.0f;
    const float K0 = (float)7.2e10, NE = (float)(-8.314e4), RG = (float)8.314;
    const float HK = (float)(6.78e4 * 7.2e10), RC = (float)(1000 * 0.239);
    const float KC = (float)((1000 * 0.239) / (1000 * 0.239 * 100));
    const float NUA = (float)(-(6.6e5 * 8.958)), DTf = (float)0.1;

    *step_count += 1; /* :396 */
    float na[2], F[2], x[4];
    for (int j = 0; j < 2; ++j) {
        na[j] = clipf(a_in[j], -1.0f, 1.0f);                 /* :399 */
        F[j] = ALO + (na[j] + 1.0f) * (AHI - ALO) / 2.0f;    /* :148-150 */
    }
    for (int j = 0; j < 4; ++j) {
        float v = SLO[j] + (s[j] + 1.0f) * (SHI[j] - SLO[j]) / 2.0f; /* :136-138 */
        x[j] = clipf(v, SLO[j], SHI[j]);                              /* :406-410 */
    }
    int bad = (x[0] != x[0]) | (x[1] != x[1]) | (x[2] != x[2]) | (x[3] != x[3]) | (F[0] != F[0]) | (F[1] != F[1]);
    if (bad) { /* :415-421 */
        *reward = -10.0f;
        *truncated = 1;
        return 1;
    }
    float C1 = x[0], T1 = x[1], C2 = x[2], T2 = x[3];
    T1 = T1 < 273.15f ? 273.15f : T1; /* :470-471 */
    T2 = T2 < 273.15f ? 273.15f : T2;
    float F1 = clipf(F[0], 1e-5f, 1e5f), F2 = clipf(F[1], 1e-5f, 1e5f); /* :472-473 */

    float a1 = clipf(NE / (RG * T1), -100.0f, 100.0f), a2 = clipf(NE / (RG * T2), -100.0f, 100.0f);
    float b1 = clipf(NUA / ((F1 * 1000.0f) * 0.239f), -100.0f, 100.0f);
    float b2 = clipf(NUA / ((F2 * 1000.0f) * 0.239f), -100.0f, 100.0f);
    float k1, k2, c1, c2;
    if (exp_mode == EXP_SHARED) {
        k1 = cstr_expf_shared(a1); k2 = cstr_expf_shared(a2);
        c1 = cstr_expf_shared(b1); c2 = cstr_expf_shared(b2);
    } else {
        k1 = expf(a1); k2 = expf(a2); c1 = expf(b1); c2 = expf(b2);
    }
    /* :479-491 */
    float dC1 = QV * (CF - C1) - (K0 * C1) * k1;
    float dT1 = (QV * (TF - T1) + ((HK * C1) / RC) * k1) + ((KC * F1) * (1.0f - c1)) * (TCF - T1);
    float dC2 = QV * (C1 - C2) - (K0 * C2) * k2;
    float dT2 = (QV * (T1 - T2) + ((HK * C2) / RC) * k2) + ((KC * F2) * (1.0f - c2)) * (TCF - T2);
    /* :493-503, then :424-428 (second clip is idempotent) */
    float nx[4] = {C1 + dC1 * DTf, T1 + dT1 * DTf, C2 + dC2 * DTf, T2 + dT2 * DTf};
    for (int j = 0; j < 4; ++j) {
        float v = clipf(clipf(nx[j], SLO[j], SHI[j]), SLO[j], SHI[j]);
        s[j] = 2.0f * (v - SLO[j]) / (SHI[j] - SLO[j]) - 1.0f; /* :131-132 */
    }
    /* compute_reward on the round-tripped state (Q12) */
    float r[4];
    for (int j = 0; j < 4; ++j) r[j] = SLO[j] + (s[j] + 1.0f) * (SHI[j] - SLO[j]) / 2.0f;
    float err = fabsf(r[2] - target);                 /* :288 */
    float n = err / (float)(0.45 - 0.05);             /* :290 */
    float n2 = sq_mode == SQ_MUL ? n * n : powf_ptr(n, 2.0f);
    float conc = -5.0f * n2 - 2.0f * n;               /* :291 */
    float tp = 0.0f;
    for (int j = 1; j < 4; j += 2) {                  /* :333-341 */
        float T = r[j];
        if (T < 280.0f) tp = tp - 0.2f * ((280.0f - T) / 280.0f);
        else if (T > 350.0f) tp = tp - 0.5f * ((T - 350.0f) / 350.0f);
    }
    *reward = 1.0f * conc + 0.5f * tp;                /* :369-377 */
    *truncated = *step_count >= MAX_STEPS;            /* :438 */
    return 0;
}

/* float64 twin ("same scheme in fp64": fp64 affine maps + _dynamics fed float64; bounds are the
 * float32 constants widened) */
static inline int step_one_f64(double s[4], const double a_in[2], int32_t *step_count, double *reward,
                               int *truncated, double target) {
    const double QV = 50.0 / 100.0, CF = 0.5, TF = 320, TCF = 370, K0 = 7.2e10, NE = -8.314e4, RG = 8.314;
    const double HK = 6.78e4 * 7.2e10, RC = 1000 * 0.239, KC = (1000 * 0.239) / (1000 * 0.239 * 100);
    const double NUA = -(6.6e5 * 8.958), DTd = 0.1;
    double slo[4], shi[4];
    for (int j = 0; j < 4; ++j) { slo[j] = (double)SLO[j]; shi[j] = (double)SHI[j]; }
    *step_count += 1;
    double na[2], F[2], x[4];
    for (int j = 0; j < 2; ++j) {
        na[j] = clipd(a_in[j], -1.0, 1.0);
        F[j] = 30.0 + (na[j] + 1.0) * (250.0 - 30.0) / 2.0;
    }
    for (int j = 0; j < 4; ++j) x[j] = clipd(slo[j] + (s[j] + 1.0) * (shi[j] - slo[j]) / 2.0, slo[j], shi[j]);
    int bad = (x[0] != x[0]) | (x[1] != x[1]) | (x[2] != x[2]) | (x[3] != x[3]) | (F[0] != F[0]) | (F[1] != F[1]);
    if (bad) { *reward = -10.0; *truncated = 1; return 1; }
    double C1 = x[0], T1 = x[1], C2 = x[2], T2 = x[3];
    T1 = T1 < 273.15 ? 273.15 : T1;
    T2 = T2 < 273.15 ? 273.15 : T2;
    double F1 = clipd(F[0], 1e-5, 1e5), F2 = clipd(F[1], 1e-5, 1e5);
    double k1 = exp(clipd(NE / (RG * T1), -100, 100)), k2 = exp(clipd(NE / (RG * T2), -100, 100));
    double c1 = exp(clipd(NUA / ((F1 * 1000) * 0.239), -100, 100)), c2 = exp(clipd(NUA / ((F2 * 1000) * 0.239), -100, 100));
    double dC1 = QV * (CF - C1) - (K0 * C1) * k1;
    double dT1 = (QV * (TF - T1) + ((HK * C1) / RC) * k1) + ((KC * F1) * (1.0 - c1)) * (TCF - T1);
    double dC2 = QV * (C1 - C2) - (K0 * C2) * k2;
    double dT2 = (QV * (T1 - T2) + ((HK * C2) / RC) * k2) + ((KC * F2) * (1.0 - c2)) * (TCF - T2);
    double nx[4] = {C1 + dC1 * DTd, T1 + dT1 * DTd, C2 + dC2 * DTd, T2 + dT2 * DTd};
    for (int j = 0; j < 4; ++j) {
        double v = clipd(nx[j], slo[j], shi[j]);
        s[j] = 2.0 * (v - slo[j]) / (shi[j] - slo[j]) - 1.0;
    }
    double r[4];
    for (int j = 0; j < 4; ++j) r[j] = slo[j] + (s[j] + 1.0) * (shi[j] - slo[j]) / 2.0;
    double err = fabs(r[2] - target), n = err / (0.45 - 0.05);
    double conc = -5.0 * (n * n) - 2.0 * n, tp = 0.0;
    for (int j = 1; j < 4; j += 2) {
        double T = r[j];
        if (T < 280) tp = tp - 0.2 * ((280 - T) / 280);
        else if (T > 350) tp = tp - 0.5 * ((T - 350) / 350);
    }
    *reward = 1.0 * conc + 0.5 * tp;
    *truncated = *step_count >= MAX_STEPS;
    return 0;
}

/* ---------------------------------------------------------------------------------------------
 * Philox4x32-10 and the reset recipe (mirror of the product's device RNG; DESIGN.md "RNG")
 *   counter = (env_id_lo, env_id_hi, episode, (stream<<8)|call)   key = (seed_lo, seed_hi)
 * ------------------------------------------------------------------------------------------- */
static inline void philox4x32_10(uint32_t c[4], uint32_t k0_, uint32_t k1_) {
    for (int i = 0; i < 10; ++i) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0_, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1_, n3 = (uint32_t)p0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0_ += 0x9E3779B9u; k1_ += 0xBB67AE85u;
    }
}

void cstr_philox_array(const uint32_t *ctr, const uint32_t *key, uint32_t *out, int64_t n) {
    for (int64_t i = 0; i < n; ++i) {
        uint32_t c[4] = {ctr[4 * i], ctr[4 * i + 1], ctr[4 * i + 2], ctr[4 * i + 3]};
        philox4x32_10(c, key[0], key[1]);
        memcpy(out + 4 * i, c, 16);
    }
}

#define STREAM_RESET 1u

static inline double u53(uint32_t hi, uint32_t lo) {
    return ((double)(hi >> 5) * 67108864.0 + (double)(lo >> 6)) / 9007199254740992.0;
}

/* 8 unit doubles for (env, episode) */
static inline void reset_uniforms(uint64_t seed, uint64_t env, uint32_t episode, double u[8]) {
    for (uint32_t call = 0; call < 4; ++call) {
        uint32_t c[4] = {(uint32_t)env, (uint32_t)(env >> 32), episode, (STREAM_RESET << 8) | call};
        philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
        u[2 * call] = u53(c[0], c[1]);
        u[2 * call + 1] = u53(c[2], c[3]);
    }
}

void cstr_oracle_reset_uniforms(uint64_t seed, int64_t env0, int64_t n, const int32_t *episode, double *u_out) {
    for (int64_t i = 0; i < n; ++i) reset_uniforms(seed, (uint64_t)(env0 + i), (uint32_t)episode[i], u_out + 8 * i);
}

/* generate_initial_state (:167-224) from 8 uniforms, all float64 */
static inline void initial_raw_random(const double u[8], double s[4]) {
    const double lo[4] = {0.05, 280.0, 0.05, 280.0}, hi[4] = {0.45, 380.0, 0.45 * 0.8, 380.0};
    for (int j = 0; j < 4; ++j) s[j] = lo[j] + (hi[j] - lo[j]) * u[j];
    for (int j = 0; j < 4; ++j) s[j] = s[j] + (-0.05 + (0.05 - (-0.05)) * u[4 + j]);
    if (s[1] < s[3]) { double t = s[1]; s[1] = s[3]; s[3] = t; }
    if (s[0] < s[2]) { double t = s[0]; s[0] = s[2]; s[2] = t; }
    for (int j = 0; j < 4; ++j) s[j] = clipd(s[j], (double)SLO[j], (double)SHI[j]);
}

/* static mode (:245-253): base += uniform(lo, hi) in place (quirk Q2), no clip */
static inline void initial_raw_static(const double u[8], double base[4], double s[4]) {
    const double lo[4] = {-0.05, -10.0, -0.05, -10.0}, hi[4] = {0.05, 10.0, 0.05, 10.0};
    for (int j = 0; j < 4; ++j) { base[j] = base[j] + (lo[j] + (hi[j] - lo[j]) * u[j]); s[j] = base[j]; }
}

static inline void normalize_raw64(const double s[4], double o[4]) {
    for (int j = 0; j < 4; ++j) o[j] = 2.0 * (s[j] - (double)SLO[j]) / ((double)SHI[j] - (double)SLO[j]) - 1.0;
}

/* ---------------------------------------------------------------------------------------------
 * exported entry points
 * ------------------------------------------------------------------------------------------- */

/* single step, no auto-reset: the pure TwoSeriesCSTREnv.step for N envs */
void cstr_oracle_step_f32(float *state, const float *action, int32_t *step_count, float *reward,
                          uint8_t *truncated, uint8_t *nan_row, int64_t n, float target, int exp_mode,
                          int sq_mode) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        int tr;
        int bad = step_one_f32(state + 4 * i, action + 2 * i, step_count + i, reward + i, &tr, target, exp_mode, sq_mode);
        truncated[i] = (uint8_t)tr;
        if (nan_row) nan_row[i] = (uint8_t)bad;
    }
}

void cstr_oracle_step_f64(double *state, const double *action, int32_t *step_count, double *reward,
                          uint8_t *truncated, uint8_t *nan_row, int64_t n, double target) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        int tr;
        int bad = step_one_f64(state + 4 * i, action + 2 * i, step_count + i, reward + i, &tr, target);
        truncated[i] = (uint8_t)tr;
        if (nan_row) nan_row[i] = (uint8_t)bad;
    }
}

/* reset of the envs flagged in `mask` (NULL = all) with the Philox recipe; init_mode 0=random 1=static */
void cstr_oracle_reset_f32(float *state, int32_t *step_count, int32_t *episode, double *static_base,
                           const uint8_t *mask, int64_t n, int64_t env0, uint64_t seed, int init_mode) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        if (mask && !mask[i]) continue;
        double u[8], s[4], o[4];
        reset_uniforms(seed, (uint64_t)(env0 + i), (uint32_t)episode[i], u);
        if (init_mode == 0) initial_raw_random(u, s); else initial_raw_static(u, static_base + 4 * i, s);
        normalize_raw64(s, o);
        for (int j = 0; j < 4; ++j) state[4 * i + j] = (float)o[j];
        step_count[i] = 0;
        episode[i] += 1;
    }
}

/* T-step tape with DummyVecEnv auto-reset semantics.  actions: (T,N,2).  Optional outputs (NULL to
 * skip): rewards (T,N), dones (T,N), obs_tape (T,N,4) = the observation RETURNED by step_wait (post
 * reset on done rows), term_tape (T,N,4) = the terminal observation.  Returns via reward_sum the
 * sum of all rewards in double (a cheap checksum for large runs). */
void cstr_oracle_tape_f32(float *state, int32_t *step_count, int32_t *episode, double *static_base,
                          const float *actions, int64_t T, int64_t n, int64_t env0, uint64_t seed,
                          int init_mode, float target, int exp_mode, int sq_mode, float *rewards,
                          uint8_t *dones, float *obs_tape, float *term_tape, double *reward_sum) {
    double total = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : total)
    for (int64_t i = 0; i < n; ++i) {
        float s[4];
        memcpy(s, state + 4 * i, 16);
        int32_t sc = step_count[i], ep = episode[i];
        for (int64_t t = 0; t < T; ++t) {
            float r; int tr;
            step_one_f32(s, actions + (t * n + i) * 2, &sc, &r, &tr, target, exp_mode, sq_mode);
            total += (double)r;
            if (rewards) rewards[t * n + i] = r;
            if (dones) dones[t * n + i] = (uint8_t)tr;
            if (term_tape) memcpy(term_tape + (t * n + i) * 4, s, 16);
            if (tr) {
                double u[8], raw[4], o[4];
                reset_uniforms(seed, (uint64_t)(env0 + i), (uint32_t)ep, u);
                if (init_mode == 0) initial_raw_random(u, raw); else initial_raw_static(u, static_base + 4 * i, raw);
                normalize_raw64(raw, o);
                for (int j = 0; j < 4; ++j) s[j] = (float)o[j];
                sc = 0; ep += 1;
            }
            if (obs_tape) memcpy(obs_tape + (t * n + i) * 4, s, 16);
        }
        memcpy(state + 4 * i, s, 16);
        step_count[i] = sc; episode[i] = ep;
    }
    if (reward_sum) *reward_sum = total;
}

void cstr_oracle_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int cstr_oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
