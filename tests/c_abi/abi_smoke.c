/* Plain-C client of libtorj_cuda.so: what a foreign host (Julia's ccall, C, Fortran) does — no Python, no torch.
 * Builds a tiny analytic equilibrium, hands the RAW arrays to torj_plasma_create_from_data, traces a 3-ray bundle with
 * a trajectory window and prints a few numbers the pytest wrapper checks.   gcc abi_smoke.c -I../../include -L... -ltorj_cuda -lm */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include "torj_cuda.h"

#define CHECK(x) do { int rc_ = (x); if (rc_) { fprintf(stderr, "FAIL %s: %s\n", #x, torj_last_error()); return 1; } } while (0)

int main(void) {
    const int nR = 65, nZ = 65, nprof = 51, ngl = 8, npsi = 40;
    const double R0 = 1.7, a = 0.6, E = 1.7, A = 0.08, B0 = 2.0;
    torj_grid g = {nR, nZ, 0.5, 2.7, -1.4, 1.4};
    double *psi = malloc(sizeof(double) * nR * nZ), *BR = malloc(sizeof(double) * nR * nZ), *BZ = malloc(sizeof(double) * nR * nZ),
           *Bp = malloc(sizeof(double) * nR * nZ);
    const double psib = A * pow((R0 + a) * (R0 + a) - R0 * R0, 2) / 4.0;
    for (int j = 0; j < nZ; ++j)
        for (int i = 0; i < nR; ++i) {
            double R = g.R_first + (g.R_last - g.R_first) * i / (nR - 1), Z = g.Z_first + (g.Z_last - g.Z_first) * j / (nZ - 1);
            int k = j * nR + i; /* R fastest */
            psi[k] = A * (pow(R * R - R0 * R0, 2) / 4.0 + R * R * Z * Z / (E * E)) / psib;
            BR[k] = -2.0 * A * R * Z / (E * E);
            BZ[k] = A * ((R * R - R0 * R0) + 2.0 * Z * Z / (E * E));
            Bp[k] = B0 * R0 / R;
        }
    double pp[51], ne[51], Te[51], vol[51];
    for (int i = 0; i < nprof; ++i) {
        pp[i] = (double)i / (nprof - 1);
        ne[i] = 3e19 * (1 - pp[i]) + 1e17;
        Te[i] = 4e3 * (1 - pp[i]) * (1 - pp[i]) + 50.0;
        vol[i] = 2 * M_PI * M_PI * R0 * a * a * E * pp[i];
    }
    /* 8-point Gauss-Legendre */
    double t[8] = {-0.9602898564975363, -0.7966664774136267, -0.5255324099163290, -0.1834346424956498,
                   0.1834346424956498, 0.5255324099163290, 0.7966664774136267, 0.9602898564975363};
    double w[8] = {0.1012285362903763, 0.2223810344533745, 0.3137066458778873, 0.3626837833783620,
                   0.3626837833783620, 0.3137066458778873, 0.2223810344533745, 0.1012285362903763};
    torj_ctx* ctx; torj_plasma* pl;
    if (torj_abi_version() != TORJ_ABI_VERSION) { fprintf(stderr, "ABI version mismatch\n"); return 1; }
    CHECK(torj_ctx_create(0, NULL, &ctx));
    CHECK(torj_abs_init(ctx, ngl, t, w));
    CHECK(torj_plasma_create_from_data(ctx, &g, psi, pp, ne, Te, nprof, BR, BZ, Bp, pp, vol, nprof, &pl));
    /* three rays: [3][n] component-major; the second points away from the plasma */
    double pos[9] = {2.5, 2.5, 2.5, 0.0, 0.0, 0.01, 0.4, 0.4, 0.38};
    double dir[9] = {-0.8660254037844387, 0.8660254037844387, -0.8660254037844387, 0, 0, 0, -0.5, 0.5, -0.5};
    double wt[3] = {0.5, 0.25, 0.25}, f = 95e9;
    int32_t mode = 1, npts[3], st[3];
    double edges[40], prof[40], dep, Pf[3], Pd[3];
    for (int j = 0; j < npsi; ++j) edges[j] = (double)j / (npsi - 1);
    const int maxp = 2 + 100 * 58; /* 100 segments x (0.5 m / 100 / dtmax = 50 steps + slack) */
    double *ts = calloc(maxp, 8), *txyz = calloc(3 * maxp, 8), *tP = calloc(maxp, 8), *tdP = calloc(maxp, 8), *tprof = calloc(npsi, 8);
    torj_counters cnt;
    torj_options opt; torj_options_default(&opt);
    CHECK(torj_trace(ctx, pl, &opt, 3, pos, dir, wt, &f, &mode, 0, 0.5, npsi, edges, 1, NULL, prof, &dep, Pf, Pd, npts, st,
                     0, 1, maxp, ts, txyz, tP, tdP, tprof, &cnt));
    printf("status %d %d %d\n", st[0], st[1], st[2]);
    printf("npts %d %d %d\n", npts[0], npts[1], npts[2]);
    printf("P_final %.12e %.12e %.12e\n", Pf[0], Pf[1], Pf[2]);
    printf("deposited %.12e\n", dep);
    printf("n_acc %lld\n", (long long)cnt.n_acc);
    printf("traj0 s_end %.12e x_end %.12e z_end %.12e\n", ts[npts[0] - 1], txyz[npts[0] - 1], txyz[2 * maxp + npts[0] - 1]);
    double sum = 0; for (int j = 0; j < npsi; ++j) sum += prof[j];
    printf("profile_sum %.12e last %.3e\n", sum, prof[npsi - 1]);
    torj_plasma_destroy(pl);
    torj_ctx_destroy(ctx);
    return 0;
}
