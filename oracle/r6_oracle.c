/*
 * r6_oracle.c — CPU restatement of the reference's 6DOF env-step path.  TEST INFRASTRUCTURE ONLY.
 *
 * Nothing on the product path may link, import or call this file: only tests/, the smoke check in
 * __graft_entry__.py and bench.py's cpu_baseline / --impl reference legs use it, as the checker
 * or the timed CPU arm.  The product (rl_rocket_6dof_b200) fails loudly without its CUDA library.
 *
 * Parity pin: tests/test_oracle_golden.py checks every function here against the fixtures in
 * tests/golden/, which were produced by RUNNING the unmodified reference (oracle/make_golden.py)
 * under numpy 2.3.5 / scipy 1.18.1 in the build container.
 *
 * Citations: "sim:" = /root/reference/my_environment/utils/simulator.py,
 *            "env:" = /root/reference/my_environment/envs/rocket_env.py,
 *            "rk:", "common:", "ivp:", "base:" = scipy/integrate/_ivp/{rk,common,ivp,base}.py (1.18.1),
 *            "rot:" = scipy/spatial/transform/_rotation_xp.py, "brentq" = scipy/optimize/Zeros/brentq.c.
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off: no implicit FMA, every float32 step of the
 * reference's mixed-precision map is an explicit float operation).
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

/* ------------------------------------------------------------------------------------------ */
/* Parameters (host fills them with the reference's own Python expressions, see params.py)     */
typedef struct {
    double dt;                /* env: timestep */
    double max_gimbal;        /* env:86  np.deg2rad(20) (float64 scalar) */
    float  max_thrust;        /* env:87  981e3, used in float32 arithmetic */
    float  beta;              /* env:355 float32(beta) */
    float  w_v_f;             /* env:401 */
    float  w_r_f;             /* env:400 */
    float  max_r_f;           /* env:400 */
    float  max_v_f;           /* env:401 */
    float  maximum_v;         /* env:384 */
    float  target_r;          /* env:385 */
    float  zero_height_tol;   /* env:383 float32(1e-3) */
    float  bounds_low[3];     /* env:129-134 */
    float  bounds_high[3];
    double normalizer[14];    /* env:106-126 */
    double alfa, eta, gamma, kappa;           /* env:343,356,367,399 */
    double att_traj_limit[3]; /* env:159 */
    double land_att_limit[3]; /* env:165 */
    double omega_lim[3];      /* env:167 */
    double waypoint;          /* env:168 */
    int32_t shaping_velocity; /* env:331/345: 0 = 'acceleration', 1 = 'velocity' */
    int32_t n_t;              /* length of t_table */
    const double *t_table;    /* t_k = round(t_{k-1}+dt, 3), sim:92 — built on the host */
} R6OParams;

typedef struct {
    double y[14];     /* sim: state (float64 after the first step; float32 IC widened before) */
    float  m0;        /* sim:42  initial mass (float32 in env mode) */
    float  v0;        /* env:651 ||IC[3:6]|| (float32), velocity shaping only */
    int32_t k;        /* steps taken in this episode (t = t_table[k]) */
    int32_t pad;
} R6OEnv;

typedef struct {
    double state[14];
    float  obs[14];
    double reward;        /* sum of terms (+ -50 if out of bounds); NOT clipped */
    double terms[7];      /* shaping, thrust_penalty, eta, attitude_constraint, goal, final_pos, final_vel */
    float  u[3];          /* de-normalised action */
    int32_t done, oob, status, nfev;
    int32_t flags[5];     /* zero_height, velocity_limit, landing_radius, attitude_limit, omega_limit */
    int32_t tgo_npos;     /* number of positive real roots of the t_go quartic (0 => reference raises) */
} R6OOut;

/* ------------------------------------------------------------------------------------------ */
/* float32 helpers: NumPy's float32 sin/cos on |x| < pi/4 and OpenBLAS sdot (§C.3) */
static float np_cosf_small(float x)
{
    float x2 = x * x, r;
    r = fmaf(0x1.98e616p-16f, x2, -0x1.6c06dcp-10f);
    r = fmaf(r, x2, 0x1.55553cp-05f);
    r = fmaf(r, x2, -0x1p-1f);
    r = fmaf(r, x2, 1.0f);
    return r;
}
static float np_sinf_small(float x)
{
    float x2 = x * x, r;
    r = fmaf(0x1.7d3bbcp-19f, x2, -0x1.a06bbap-13f);
    r = fmaf(r, x2, 0x1.11119ap-07f);
    r = fmaf(r, x2, -0x1.555556p-03f);
    r = fmaf(r, x2, 0.0f);
    r = fmaf(r, x, x);
    return r;
}
static float sdot3(const float *a, const float *b)
{
    double acc = (double)(float)(a[0] * b[0]);
    acc += (double)(float)(a[1] * b[1]);
    acc += (double)(float)(a[2] * b[2]);
    return (float)acc;
}
static float snrm3(const float *a) { return sqrtf(sdot3(a, a)); }

/* ------------------------------------------------------------------------------------------ */
/* Per-episode / per-step constants consumed by the RHS (all float64 once built)               */
typedef struct {
    double J0, J1, Ji0, Ji1;  /* sim:45-50 */
    double Tb[3];             /* sim:167-175 thrust in body frame */
    double dm;                /* sim:140-141 */
} StepConst;

/* env mode: float32 rules of SURVEY §A.1 (NumPy-2 promotion) */
static void consts_env_mode(float m0, const float u[3], StepConst *c)
{
    const float rb = 3.66f / 2;   /* python floats are "weak": all float32 below */
    /* sim:46  .5*m*base_radius**2 ; base_radius**2 is a python float (float64) product first */
    const double rb2 = (3.66 / 2) * (3.66 / 2);
    const double l2 = 40.0 * 40.0 + 3 * rb2;
    (void)rb;
    float J0 = (float)(0.5f * m0) * (float)rb2;
    float J1 = (float)((float)(1.0 / 12) * m0) * (float)l2;
    c->J0 = J0; c->J1 = J1;
    c->Ji0 = (double)(1.0f / J0);   /* sim:50 np.linalg.inv on a float32 diagonal */
    c->Ji1 = (double)(1.0f / J1);
    float cy = np_cosf_small(u[0]), cz = np_cosf_small(u[1]);
    float sy = np_sinf_small(u[0]), sz = np_sinf_small(u[1]);
    float r00 = cy * cz, r10 = sy * cz;         /* sim:210-212 float32 products */
    double T = (double)u[2];
    c->Tb[0] = (double)r00 * T;                  /* sim:174 float64 matrix @ [T,0,0] */
    c->Tb[1] = (double)r10 * T;
    c->Tb[2] = (double)sz * T;
    float g0isp = (float)(9.81 * 360);           /* sim:141 python float product, then weak -> f32 */
    c->dm = (double)((-u[2]) / g0isp);
}

/* raw simulator mode (python-list IC and control => everything float64), test_6DOF_simulator.py */
static void consts_raw_mode(double m0, const double u[3], StepConst *c)
{
    const double rb = 3.66 / 2;
    c->J0 = .5 * m0 * (rb * rb);
    c->J1 = 1.0 / 12 * m0 * (40.0 * 40.0 + 3 * (rb * rb));
    c->Ji0 = 1.0 / c->J0;
    c->Ji1 = 1.0 / c->J1;
    double cy = cos(u[0]), cz = cos(u[1]), sy = sin(u[0]), sz = sin(u[1]);
    c->Tb[0] = (cy * cz) * u[2];
    c->Tb[1] = (sy * cz) * u[2];
    c->Tb[2] = sz * u[2];
    c->dm = -u[2] / (9.81 * 360);
}

/* rot: as_matrix of the normalised (x,y,z,w) quaternion; sim:189-199 */
static void quat_to_matrix(const double q[4] /* leading scalar */, double R[3][3])
{
    double n = sqrt(q[1] * q[1] + q[2] * q[2] + q[3] * q[3] + q[0] * q[0]);
    double x = q[1] / n, y = q[2] / n, z = q[3] / n, w = q[0] / n;
    double x2 = x * x, y2 = y * y, z2 = z * z, w2 = w * w;
    double xy = x * y, zw = z * w, xz = x * z, yw = y * w, yz = y * z, xw = x * w;
    R[0][0] = x2 - y2 - z2 + w2; R[0][1] = 2 * (xy - zw);      R[0][2] = 2 * (xz + yw);
    R[1][0] = 2 * (xy + zw);     R[1][1] = -x2 + y2 - z2 + w2; R[1][2] = 2 * (yz - xw);
    R[2][0] = 2 * (xz - yw);     R[2][1] = 2 * (yz + xw);      R[2][2] = -x2 - y2 + z2 + w2;
}

/* sim:106-143 */
static void rhs(const StepConst *c, const double y[14], double f[14])
{
    /* sim:145-150 */
    const double expo = 1 + 9.81 * 0.0289644 / 8.3144598 / (-0.0065);
    double rho = 1.225 * pow(288.15 / (288.15 + (y[0] - 0) * (-0.0065)), expo);
    double R[3][3];
    quat_to_matrix(&y[6], R);
    const double *v = &y[3], *w = &y[10];
    /* sim:216-219 */
    double vb[3], A[3], Fb[3], F[3];
    for (int i = 0; i < 3; i++) vb[i] = R[0][i] * v[0] + R[1][i] * v[1] + R[2][i] * v[2];
    double vn = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    const double rb = 3.66 / 2;
    const double S_ref = M_PI * (rb * rb);
    double ca = (((-.5 * rho) * vn) * S_ref) * 0.82;
    for (int i = 0; i < 3; i++) A[i] = ca * vb[i];
    /* sim:156-165 */
    for (int i = 0; i < 3; i++) Fb[i] = c->Tb[i] + A[i];
    for (int i = 0; i < 3; i++) F[i] = R[i][0] * Fb[0] + R[i][1] * Fb[1] + R[i][2] * Fb[2];
    double im = 1 / y[13];
    f[0] = v[0]; f[1] = v[1]; f[2] = v[2];                      /* sim:129 */
    f[3] = im * F[0] + (-9.81); f[4] = im * F[1] + 0; f[5] = im * F[2] + 0;   /* sim:130 */
    /* sim:136, 221-229 — the UN-normalised quaternion */
    const double *q = &y[6];
    f[6] = 0.5 * (-w[0] * q[1] - w[1] * q[2] - w[2] * q[3]);
    f[7] = 0.5 * (w[0] * q[0] + w[2] * q[2] - w[1] * q[3]);
    f[8] = 0.5 * (w[1] * q[0] - w[2] * q[1] + w[0] * q[3]);
    f[9] = 0.5 * (w[2] * q[0] + w[1] * q[1] - w[0] * q[2]);
    /* sim:232-244: r_T x T + r_cp x A with r_T = [-15,0,0], r_cp = [5,0,0] */
    double tau[3] = { 0.0, 15 * c->Tb[2] - 5 * A[2], -15 * c->Tb[1] + 5 * A[1] };
    /* sim:137 */
    double Jw[3] = { c->J0 * w[0], c->J1 * w[1], c->J1 * w[2] };
    double cr[3] = { w[1] * Jw[2] - w[2] * Jw[1], w[2] * Jw[0] - w[0] * Jw[2], w[0] * Jw[1] - w[1] * Jw[0] };
    f[10] = c->Ji0 * (tau[0] - cr[0]);
    f[11] = c->Ji1 * (tau[1] - cr[1]);
    f[12] = c->Ji1 * (tau[2] - cr[2]);
    f[13] = c->dm;                                              /* sim:140-141 */
}

/* ------------------------------------------------------------------------------------------ */
/* rk:541-565 Dormand–Prince tableau                                                          */
static const double A_[6][5] = {
    { 0, 0, 0, 0, 0 },
    { 1.0 / 5, 0, 0, 0, 0 },
    { 3.0 / 40, 9.0 / 40, 0, 0, 0 },
    { 44.0 / 45, -56.0 / 15, 32.0 / 9, 0, 0 },
    { 19372.0 / 6561, -25360.0 / 2187, 64448.0 / 6561, -212.0 / 729, 0 },
    { 9017.0 / 3168, -355.0 / 33, 46732.0 / 5247, 49.0 / 176, -5103.0 / 18656 } };
static const double B_[6] = { 35.0 / 384, 0, 500.0 / 1113, 125.0 / 192, -2187.0 / 6784, 11.0 / 84 };
static const double E_[7] = { -71.0 / 57600, 0, 71.0 / 16695, -71.0 / 1920, 17253.0 / 339200, -22.0 / 525, 1.0 / 40 };
static const double P_[7][4] = {
    { 1, -8048581381.0 / 2820520608, 8663915743.0 / 2820520608, -12715105075.0 / 11282082432 },
    { 0, 0, 0, 0 },
    { 0, 131558114200.0 / 32700410799, -68118460800.0 / 10900136933, 87487479700.0 / 32700410799 },
    { 0, -1754552775.0 / 470086768, 14199869525.0 / 1410260304, -10690763975.0 / 1880347072 },
    { 0, 127303824393.0 / 49829197408, -318862633887.0 / 49829197408, 701980252875.0 / 199316789632 },
    { 0, -282668133.0 / 205662961, 2019193451.0 / 616988883, -1453857185.0 / 822651844 },
    { 0, 40617522.0 / 29380423, -110615467.0 / 29380423, 69997945.0 / 29380423 } };

/* common:63-65 */
static double rms14(const double *x)
{
    double s = 0;
    for (int i = 0; i < 14; i++) s += x[i] * x[i];
    return sqrt(s) / sqrt(14.0);
}

/* rk:715-737 — evaluate the quartic interpolant */
typedef struct { double t_old, h, y_old[14], Q[14][4]; } Dense;
static void dense_eval(const Dense *d, double t, double *y, int only_first)
{
    double x = (t - d->t_old) / d->h;
    double p[4]; p[0] = x; p[1] = p[0] * x; p[2] = p[1] * x; p[3] = p[2] * x;
    int n = only_first ? 1 : 14;
    for (int i = 0; i < n; i++) {
        double s = d->Q[i][0] * p[0] + d->Q[i][1] * p[1] + d->Q[i][2] * p[2] + d->Q[i][3] * p[3];
        y[i] = d->h * s + d->y_old[i];
    }
}
static double event_fun(const Dense *d, double t) { double g; dense_eval(d, t, &g, 1); return g; }

/* brentq.c (scipy/optimize/Zeros) with xtol = rtol = 4*eps, maxiter = 100; ivp:52-77 */
static double brentq_event(const Dense *d, double xa, double xb)
{
    const double xtol = 4 * DBL_EPSILON, rtol = 4 * DBL_EPSILON;
    double xpre = xa, xcur = xb, xblk = 0, fblk = 0, spre = 0, scur = 0, sbis, delta, stry, dpre, dblk;
    double fpre = event_fun(d, xpre), fcur = event_fun(d, xcur);
    if (fpre == 0) return xpre;
    if (fcur == 0) return xcur;
    if (signbit(fpre) == signbit(fcur)) return NAN;   /* scipy raises ValueError */
    for (int i = 0; i < 100; i++) {
        if (fpre != 0 && fcur != 0 && (signbit(fpre) != signbit(fcur))) {
            xblk = xpre; fblk = fpre; spre = scur = xcur - xpre;
        }
        if (fabs(fblk) < fabs(fcur)) {
            xpre = xcur; xcur = xblk; xblk = xpre;
            fpre = fcur; fcur = fblk; fblk = fpre;
        }
        delta = (xtol + rtol * fabs(xcur)) / 2;
        sbis = (xblk - xcur) / 2;
        if (fcur == 0 || fabs(sbis) < delta) return xcur;
        if (fabs(spre) > delta && fabs(fcur) < fabs(fpre)) {
            if (xpre == xblk) {
                stry = -fcur * (xcur - xpre) / (fcur - fpre);
            } else {
                dpre = (fpre - fcur) / (xpre - xcur);
                dblk = (fblk - fcur) / (xblk - xcur);
                stry = -fcur * (fblk * dblk - fpre * dpre) / (dblk * dpre * (fblk - fpre));
            }
            double lim = fmin(fabs(spre), 3 * fabs(sbis) - delta);
            if (2 * fabs(stry) < lim) { spre = scur; scur = stry; }
            else { spre = sbis; scur = sbis; }
        } else { spre = sbis; scur = sbis; }
        xpre = xcur; fpre = fcur;
        if (fabs(scur) > delta) xcur += scur;
        else xcur += (sbis > 0 ? delta : -delta);
        fcur = event_fun(d, xcur);
    }
    return xcur;
}

/*
 * One solve_ivp(fun, [t, t+dt], y0, events=height) call with all defaults (sim:83-88):
 * RK45, rtol 1e-3, atol 1e-6, fresh initial step, terminal event on y[0] (direction 0).
 * Returns the scipy status (0 finished, 1 event, -1 step too small); y is the last column.
 */
static int integrate(const StepConst *c, double y[14], double t, double dt, int *nfev_out)
{
    const double rtol = 1e-3, atol = 1e-6;
    const double t_bound = t + dt;
    double f[14], sc[14], tmp[14], y1[14], f1[14];
    int nfev = 0;
    rhs(c, y, f); nfev++;                                     /* rk:96 */
    /* common:68-134 select_initial_step, order = 4 */
    double h_abs;
    {
        double L = fabs(t_bound - t);
        if (L == 0.0) h_abs = 0.0;
        else {
            for (int i = 0; i < 14; i++) sc[i] = atol + fabs(y[i]) * rtol;
            for (int i = 0; i < 14; i++) tmp[i] = y[i] / sc[i];
            double d0 = rms14(tmp);
            for (int i = 0; i < 14; i++) tmp[i] = f[i] / sc[i];
            double d1 = rms14(tmp);
            double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
            h0 = fmin(h0, L);
            for (int i = 0; i < 14; i++) y1[i] = y[i] + h0 * 1.0 * f[i];
            rhs(c, y1, f1); nfev++;
            for (int i = 0; i < 14; i++) tmp[i] = (f1[i] - f[i]) / sc[i];
            double d2 = rms14(tmp) / h0;
            double h1 = (d1 <= 1e-15 && d2 <= 1e-15) ? fmax(1e-6, h0 * 1e-3)
                                                       : pow(0.01 / fmax(d1, d2), 1.0 / 5);
            h_abs = fmin(fmin(100 * h0, h1), L);
        }
    }
    double g = y[0];                                          /* ivp: g = event(t0, y0) */
    int status = -2;
    double K[7][14];
    while (status == -2) {
        if (t == t_bound) { status = 0; break; }              /* base:188-193 (cannot happen for dt>0) */
        /* rk:111-176 _step_impl */
        double min_step = 10 * fabs(nextafter(t, INFINITY) - t);
        if (h_abs < min_step) h_abs = min_step;
        int rejected = 0, failed = 0;
        double t_new, h, y_new[14], f_new[14];
        for (;;) {
            if (h_abs < min_step) { failed = 1; break; }
            h = h_abs;
            t_new = t + h;
            if (t_new - t_bound > 0) t_new = t_bound;
            h = t_new - t;
            h_abs = fabs(h);
            /* rk:14-71 rk_step */
            memcpy(K[0], f, sizeof f);
            for (int s = 1; s < 6; s++) {
                for (int i = 0; i < 14; i++) {
                    double dy = 0;
                    for (int j = 0; j < s; j++) dy += K[j][i] * A_[s][j];
                    tmp[i] = y[i] + dy * h;
                }
                rhs(c, tmp, K[s]); nfev++;
            }
            for (int i = 0; i < 14; i++) {
                double dy = 0;
                for (int j = 0; j < 6; j++) dy += K[j][i] * B_[j];
                y_new[i] = y[i] + h * dy;
            }
            rhs(c, y_new, f_new); nfev++;
            memcpy(K[6], f_new, sizeof f_new);
            for (int i = 0; i < 14; i++) {
                double e = 0;
                for (int j = 0; j < 7; j++) e += K[j][i] * E_[j];
                double scale = atol + fmax(fabs(y[i]), fabs(y_new[i])) * rtol;
                tmp[i] = (e * h) / scale;
            }
            double err = rms14(tmp);
            if (err < 1) {
                double factor = (err == 0) ? 10.0 : fmin(10.0, 0.9 * pow(err, -0.2));
                if (rejected) factor = fmin(1.0, factor);
                h_abs *= factor;
                break;
            }
            h_abs *= fmax(0.2, 0.9 * pow(err, -0.2));
            rejected = 1;
        }
        if (failed) { status = -1; break; }                   /* ivp: 'failed' => -1, y = last stored */
        double t_old = t, y_old[14];
        memcpy(y_old, y, sizeof y_old);
        t = t_new; memcpy(y, y_new, sizeof y_new); memcpy(f, f_new, sizeof f_new);
        if (t - t_bound >= 0) status = 0;                     /* base:203-206 */
        /* ivp:134-158, 678-699 — event on y[0], direction 0, terminal */
        double g_new = y[0];
        if ((g <= 0 && g_new >= 0) || (g >= 0 && g_new <= 0)) {
            Dense d; d.t_old = t_old; d.h = t - t_old; memcpy(d.y_old, y_old, sizeof y_old);
            for (int i = 0; i < 14; i++)
                for (int m = 0; m < 4; m++) {
                    double s = 0;
                    for (int j = 0; j < 7; j++) s += K[j][i] * P_[j][m];
                    d.Q[i][m] = s;
                }
            double te = brentq_event(&d, t_old, t);
            dense_eval(&d, te, y, 0);
            status = 1;
        }
        g = g_new;
    }
    *nfev_out = nfev;
    return status;
}

static void normalize_quat64(double y[14])    /* sim:97,153-154 */
{
    double n = sqrt(y[6] * y[6] + y[7] * y[7] + y[8] * y[8] + y[9] * y[9]);
    for (int i = 6; i < 10; i++) y[i] /= n;
}

/* Raw Simulator6DOF.step (python-list inputs): y in/out, returns status. */
int r6o_sim_step_raw(double y[14], const double u[3], double m0, double t, double dt, int *nfev)
{
    StepConst c;
    consts_raw_mode(m0, u, &c);
    int st = integrate(&c, y, t, dt, nfev);
    normalize_quat64(y);
    return st;
}

/* ------------------------------------------------------------------------------------------ */
/* t_go: largest positive real root of c0 t^4 + c2 t^2 + c3 t + c4 (env:528-546; np.roots picks
 * the first eigenvalue with imag == 0 and real > 0, which is the largest positive real root —
 * SURVEY §A.4, pinned by tests/golden/units.npz).  Method here: Aberth–Ehrlich on all four complex
 * roots, then Newton polish of the real ones on the real axis.  (The CUDA kernel uses a different,
 * bracketing method; the two are checked against each other and against np.roots.)               */
typedef struct { double re, im; } cplx;
static cplx cmul(cplx a, cplx b) { cplx r = { a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re }; return r; }
static cplx cdiv(cplx a, cplx b)
{
    double d = b.re * b.re + b.im * b.im;
    cplx r = { (a.re * b.re + a.im * b.im) / d, (a.im * b.re - a.re * b.im) / d };
    return r;
}
double r6o_tgo(double c0, double c2, double c3, double c4, int *npos_out)
{
    /* monic coefficients */
    double a2 = c2 / c0, a1 = c3 / c0, a0 = c4 / c0;
    double bound = 1 + fmax(fabs(a2), fmax(fabs(a1), fabs(a0)));
    double rad = fmin(bound, 2 * fmax(sqrt(fabs(a2)), fmax(cbrt(fabs(a1)), sqrt(sqrt(fabs(a0))))));
    if (!(rad > 0)) { if (npos_out) *npos_out = 0; return NAN; }
    cplx z[4];
    for (int k = 0; k < 4; k++) {
        double th = 0.4 + 2 * M_PI * k / 4;
        z[k].re = 0.7 * rad * cos(th); z[k].im = 0.7 * rad * sin(th);
    }
    for (int it = 0; it < 200; it++) {
        double moved = 0;
        for (int k = 0; k < 4; k++) {
            cplx x = z[k], x2 = cmul(x, x);
            cplx p = cmul(x2, x2);
            p.re += a2 * x2.re + a1 * x.re + a0; p.im += a2 * x2.im + a1 * x.im;
            cplx dp = cmul(x2, x); dp.re = 4 * dp.re + 2 * a2 * x.re + a1; dp.im = 4 * dp.im + 2 * a2 * x.im;
            cplx w = cdiv(p, dp);
            cplx s = { 0, 0 };
            for (int j = 0; j < 4; j++) if (j != k) {
                cplx dif = { x.re - z[j].re, x.im - z[j].im }, one = { 1, 0 };
                cplx inv = cdiv(one, dif); s.re += inv.re; s.im += inv.im;
            }
            cplx den = cmul(w, s); den.re = 1 - den.re; den.im = -den.im;
            cplx step = cdiv(w, den);
            z[k].re -= step.re; z[k].im -= step.im;
            moved = fmax(moved, hypot(step.re, step.im) / fmax(hypot(z[k].re, z[k].im), 1e-300));
        }
        if (moved < 1e-15) break;
    }
    double best = NAN; int npos = 0;
    for (int k = 0; k < 4; k++) {
        if (fabs(z[k].im) > 1e-6 * fmax(1.0, fabs(z[k].re))) continue;
        double x = z[k].re;
        for (int it = 0; it < 8; it++) {          /* real-axis Newton polish */
            double x2 = x * x, f = (x2 + a2) * x2 + a1 * x + a0, df = (4 * x2 + 2 * a2) * x + a1;
            if (df == 0) break;
            double xn = x - f / df;
            if (xn == x) break;
            x = xn;
        }
        /* accept as real only if the polynomial changes sign around x (rejects conjugate pairs
         * that merely sit close to the axis) */
        double e = 64 * DBL_EPSILON * fmax(fabs(x), 1e-300), x2, fl, fr;
        int ok = 0;
        for (int tr = 0; tr < 20 && !ok; tr++, e *= 4) {
            double xl = x - e, xr = x + e;
            x2 = xl * xl; fl = (x2 + a2) * x2 + a1 * xl + a0;
            x2 = xr * xr; fr = (x2 + a2) * x2 + a1 * xr + a0;
            if ((fl < 0) != (fr < 0)) ok = 1;
            if (e > 1e-7 * fmax(fabs(x), 1.0)) break;
        }
        if (!ok || !(x > 0)) continue;
        npos++;
        if (!(best >= x)) best = x;
    }
    if (npos_out) *npos_out = npos;
    return best;
}

/* ------------------------------------------------------------------------------------------ */
static double wrap_pi(double a)      /* rot: (angles + pi) % (2 pi) - pi, python modulo */
{
    double m = fmod(a + M_PI, 2 * M_PI);
    if (m < 0) m += 2 * M_PI;
    return m - M_PI;
}
/* rot:365-401,1052-1111 as_euler("zyx") (extrinsic) of a leading-scalar quaternion, float64 */
void r6o_euler_zyx(const double q_in[4], double e[3])
{
    double n = sqrt(q_in[1] * q_in[1] + q_in[2] * q_in[2] + q_in[3] * q_in[3] + q_in[0] * q_in[0]);
    double w = q_in[0] / n, x = q_in[1] / n, y = q_in[2] / n, z = q_in[3] / n;
    double a = w - y, b = z - x, c = y + w, d = -x - z;
    double hs = atan2(b, a), hd = atan2(d, c);
    double a1 = 2 * atan2(hypot(c, d), hypot(a, b));
    int case1 = fabs(a1) <= 1e-7, case2 = fabs(a1 - M_PI) <= 1e-7;
    double e0, e2;
    if (!(case1 || case2)) { e0 = hs - hd; e2 = -(hs + hd); }
    else { e0 = case1 ? 2 * hs : -2 * hd; e2 = -0.0; }
    e[0] = wrap_pi(e0); e[1] = wrap_pi(a1 - M_PI / 2); e[2] = wrap_pi(e2);
}

/* env:509-521 for a float32 action array */
void r6o_denormalize_action(const R6OParams *p, const float a[3], float u[3])
{
    u[0] = (float)((double)a[0] * p->max_gimbal);
    u[1] = (float)((double)a[1] * p->max_gimbal);
    u[2] = (float)((float)(a[2] + 1.0f) / 2.0f) * p->max_thrust;
}

/* env:180-199 given the already-sampled float32 IC: float32 quaternion normalisation (:190) */
void r6o_reset_from_sample(const R6OParams *p, const float sample[14], R6OEnv *e, float obs[14], float ic_out[14])
{
    float ic[14];
    memcpy(ic, sample, sizeof ic);
    double acc = 0;
    for (int i = 6; i < 10; i++) acc += (double)(float)(ic[i] * ic[i]);
    float n = sqrtf((float)acc);
    for (int i = 6; i < 10; i++) ic[i] = ic[i] / n;
    for (int i = 0; i < 14; i++) e->y[i] = (double)ic[i];
    e->m0 = ic[13];
    e->v0 = snrm3(&ic[3]);
    e->k = 0; e->pad = 0;
    if (obs) for (int i = 0; i < 14; i++) obs[i] = (float)((double)ic[i] / p->normalizer[i]);   /* env:503-504 */
    if (ic_out) memcpy(ic_out, ic, sizeof ic);
}

/* env:201-231 one Rocket6DOF.step */
void r6o_env_step(const R6OParams *p, R6OEnv *e, const float a[3], R6OOut *o)
{
    float u[3];
    r6o_denormalize_action(p, a, u);
    StepConst c;
    consts_env_mode(e->m0, u, &c);
    int kk = e->k < p->n_t ? e->k : p->n_t - 1;
    double t = p->t_table[kk];
    int nfev = 0;
    int status = integrate(&c, e->y, t, p->dt, &nfev);
    normalize_quat64(e->y);
    e->k += 1;
    const double *S = e->y;
    float s[14];
    for (int i = 0; i < 14; i++) s[i] = (float)S[i];           /* env:206 */
    const float *r = &s[0], *v = &s[3];
    float m = s[13];
    /* env:591-593 */
    int oob = !(r[0] >= p->bounds_low[0] && r[0] <= p->bounds_high[0] &&
                r[1] >= p->bounds_low[1] && r[1] <= p->bounds_high[1] &&
                r[2] >= p->bounds_low[2] && r[2] <= p->bounds_high[2]);
    int done = (status != 0) || oob;                           /* env:213 */
    float vn = snrm3(v), rn = snrm3(r);
    /* Euler angles of the float32-cast quaternion (env:209-210, 368, 378) */
    double q32[4] = { (double)s[6], (double)s[7], (double)s[8], (double)s[9] }, eul[3];
    r6o_euler_zyx(q32, eul);
    double shaping;
    int npos = 1;
    if (!p->shaping_velocity) {
        /* env:526-566 */
        double c0 = (-9.81) * (-9.81);
        float c2 = -4 * (vn * vn), c3 = -24 * sdot3(r, v), c4 = -36 * (rn * rn);
        double tgo = r6o_tgo(c0, (double)c2, (double)c3, (double)c4, &npos);
        double tg2 = tgo * tgo;
        double qv[3];
        const double g[3] = { -9.81, 0, 0 };
        for (int i = 0; i < 3; i++) {
            float m6r = -6 * r[i], f4v = 4 * v[i];
            qv[i] = (double)m6r / tg2 - (double)f4v / tgo - g[i];
        }
        float U = p->max_thrust / m;
        double qn = sqrt(qv[0] * qv[0] + qv[1] * qv[1] + qv[2] * qv[2]);
        double at[3];
        if (qn <= (double)U) { at[0] = qv[0]; at[1] = qv[1]; at[2] = qv[2]; }
        else for (int i = 0; i < 3; i++) at[i] = qv[i] * (double)U / qn;
        /* env:339-343, sim:177-186 — float64 post-step quaternion, current control */
        double R[3][3], d[3];
        quat_to_matrix(&S[6], R);
        for (int i = 0; i < 3; i++) {
            double tv = R[i][0] * c.Tb[0] + R[i][1] * c.Tb[1] + R[i][2] * c.Tb[2];
            d[i] = tv / (double)m - at[i];
        }
        shaping = p->alfa * sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    } else {
        /* env:646-674 + :349 */
        double r_hat[3], v_hat[3], tau;
        if ((double)r[0] > p->waypoint) {
            r_hat[0] = (double)r[0] - p->waypoint; r_hat[1] = r[1]; r_hat[2] = r[2];
            v_hat[0] = (double)v[0] - (-2); v_hat[1] = v[1]; v_hat[2] = v[2];
            tau = 20;
        } else {
            r_hat[0] = (double)(float)(r[0] + 1.0f); r_hat[1] = 0; r_hat[2] = 0;
            v_hat[0] = (double)v[0] - (-1); v_hat[1] = v[1]; v_hat[2] = v[2];
            tau = 100;
        }
        double rh = sqrt(r_hat[0] * r_hat[0] + r_hat[1] * r_hat[1] + r_hat[2] * r_hat[2]);
        double vh = sqrt(v_hat[0] * v_hat[0] + v_hat[1] * v_hat[1] + v_hat[2] * v_hat[2]);
        double tgo = rh / vh, k = 1 - exp(-tgo / tau), den = fmax(1e-3, rh), d[3];
        for (int i = 0; i < 3; i++) {
            double vt = ((double)(-e->v0) * (r_hat[i] / den)) * k;
            d[i] = (double)v[i] - vt;
        }
        shaping = p->alfa * sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    }
    /* env:353-359 */
    float pen = p->beta * u[2];
    int att_viol = fabs(eul[0]) > p->att_traj_limit[0] || fabs(eul[1]) > p->att_traj_limit[1] ||
                   fabs(eul[2]) > p->att_traj_limit[2];
    double att = att_viol ? p->gamma : 0.0;
    /* env:382-390 */
    int fl[5];
    fl[0] = s[0] <= p->zero_height_tol;
    fl[1] = vn < p->maximum_v;
    fl[2] = rn < p->target_r;
    fl[3] = fabs(eul[0]) < p->land_att_limit[0] || fabs(eul[1]) < p->land_att_limit[1] ||
            fabs(eul[2]) < p->land_att_limit[2];
    fl[4] = fabs((double)s[10]) < p->omega_lim[0] || fabs((double)s[11]) < p->omega_lim[1] ||
            fabs((double)s[12]) < p->omega_lim[2];
    double goal = (fl[0] && fl[1] && fl[2] && fl[3] && fl[4]) ? p->kappa : 0.0;
    /* env:400-401 */
    float dr = p->max_r_f - rn;
    double final_pos = dr > 0 ? (double)(dr * p->w_r_f) : 0.0;
    float dv = p->max_v_f - vn;
    double final_vel = (rn < p->max_r_f && fl[0]) ? (dv > 0 ? (double)(dv * p->w_v_f) : 0.0) : 0.0;
    double reward = 0;
    reward += shaping; reward += (double)pen; reward += p->eta; reward += att; reward += goal;
    reward += final_pos; reward += final_vel;                  /* env:362 */
    if (oob) reward += -50;                                    /* env:228-229 */
    memcpy(o->state, S, sizeof o->state);
    for (int i = 0; i < 14; i++) o->obs[i] = (float)(S[i] / p->normalizer[i]);   /* env:503-504 */
    o->reward = reward;
    o->terms[0] = shaping; o->terms[1] = pen; o->terms[2] = p->eta; o->terms[3] = att;
    o->terms[4] = goal; o->terms[5] = final_pos; o->terms[6] = final_vel;
    memcpy(o->u, u, sizeof u);
    o->done = done; o->oob = oob; o->status = status; o->nfev = nfev;
    memcpy(o->flags, fl, sizeof fl);
    o->tgo_npos = npos;
}

typedef struct {
    const R6OParams *p; R6OEnv *envs; const float *actions; R6OOut *outs; int64_t lo, hi;
} BatchJob;
static void *batch_worker(void *arg)
{
    BatchJob *j = (BatchJob *)arg;
    for (int64_t i = j->lo; i < j->hi; i++) r6o_env_step(j->p, &j->envs[i], &j->actions[3 * i], &j->outs[i]);
    return NULL;
}
/* envs are independent: contiguous index ranges per POSIX thread (no OpenMP runtime in the image) */
void r6o_env_step_batch(const R6OParams *p, R6OEnv *envs, int64_t n, const float *actions /* [n][3] */,
                        R6OOut *outs, int nthreads)
{
    if (nthreads > n) nthreads = (int)n;
    if (nthreads <= 1) {
        for (int64_t i = 0; i < n; i++) r6o_env_step(p, &envs[i], &actions[3 * i], &outs[i]);
        return;
    }
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)nthreads);
    BatchJob *jobs = (BatchJob *)malloc(sizeof(BatchJob) * (size_t)nthreads);
    for (int k = 0; k < nthreads; k++) {
        jobs[k] = (BatchJob){ p, envs, actions, outs, n * k / nthreads, n * (k + 1) / nthreads };
        pthread_create(&th[k], NULL, batch_worker, &jobs[k]);
    }
    for (int k = 0; k < nthreads; k++) pthread_join(th[k], NULL);
    free(th); free(jobs);
}

/* sizes for the ctypes layout check */
int r6o_sizeof_params(void) { return (int)sizeof(R6OParams); }
int r6o_sizeof_env(void) { return (int)sizeof(R6OEnv); }
int r6o_sizeof_out(void) { return (int)sizeof(R6OOut); }

/* unit hooks for the golden tests */
void r6o_step_consts(float m0, const float u[3], double out[8])
{
    StepConst c; consts_env_mode(m0, u, &c);
    out[0] = c.J0; out[1] = c.J1; out[2] = c.Ji0; out[3] = c.Ji1;
    out[4] = c.Tb[0]; out[5] = c.Tb[1]; out[6] = c.Tb[2]; out[7] = c.dm;
}
