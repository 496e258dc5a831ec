// CPU differential test: the restated ARMS (hydra_b200/csrc/arms.cuh, the same source the device compiles)
// against the reference's own ARMS object code (oracle/_ref/libarms_ref.so), fed the same rand() integers.
#include <dlfcn.h>
#include <sys/wait.h>
#include <unistd.h>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#include "../../hydra_b200/csrc/bayesw_math.cuh"

typedef int (*ref_arms_t)(const double *, int, double, double, double (*)(double, void *), void *, int (*)(void *), void *, double *, int *);

struct Lcg { uint64_t s; long calls; };
static int lcg_rand(void *d) {  // 31-bit integers like glibc rand()
    Lcg *g = (Lcg *)d;
    g->s = g->s * 6364136223846793005ull + 1442695040888963407ull;
    g->calls++;
    return (int)((g->s >> 33) & 0x7FFFFFFFu);
}
static double dens_cb(double x, void *d) { return (*(hb::BwBetaDens *)d)(x); }

int main(int argc, char **argv) {
    void *h = dlopen(argc > 1 ? argv[1] : "oracle/_ref/libarms_ref.so", RTLD_NOW);
    if (!h) { printf("SKIP %s\n", dlerror()); return 77; }
    ref_arms_t ref = (ref_arms_t)dlsym(h, "ho_ref_arms");
    int bad = 0, n = 0, fail_both = 0;
    long evals = 0;
    Lcg par{12345, 0};
    auto U = [&]() { return (lcg_rand(&par) + 0.5) / 2147483648.0; };
    for (int t = 0; t < (argc > 2 ? atoi(argv[2]) : 4000); t++) {
        hb::BwMarker m;
        m.alpha = 1.0 + 12.0 * U();
        m.sigmaG = 0.001 + 0.05 * U();
        const double p = 0.001 + 0.499 * U();
        m.mean = 2 * p; m.sd = sqrt(2 * p * (1 - p)); m.mean_sd_ratio = m.mean / m.sd;
        const double N = 1000 + 400000 * U();
        m.vi_1 = N * 2 * p * (1 - p) * (0.5 + U()); m.vi_2 = N * p * p * (0.5 + U()); m.vi_0 = N * (1 - p) * (1 - p) * (0.5 + U());
        m.vi_sum = m.vi_0 + m.vi_1 + m.vi_2;
        m.sum_failure = (U() - 0.5) * sqrt(N) * 2;
        const double Ck = pow(10.0, -4 + 3 * U()), sumSG = m.sigmaG * (1 + U());
        const double bold = (U() < 0.5) ? 0.0 : (U() - 0.5) * 0.05;
        const uint64_t seed = 777 + t;
        // reference
        Lcg g1{seed, 0};
        hb::BwBetaDens d{m, Ck};
        const double L = 2 * sqrt(sumSG * Ck);
        const double xinit[4] = {bold - L / 10, bold, bold + L / 20, bold + L / 10};
        double xr = 0; int nev = 0;
        int e1 = 0;
        {   // the reference exit()s on internal imprecision (EXIT1..EXIT6): run it in a child
            int fd[2];
            if (pipe(fd)) return 2;
            fflush(stdout);
            const pid_t pid = fork();
            if (pid == 0) {
                close(fd[0]);
                if (!freopen("/dev/null", "w", stdout)) _exit(3);
                double x = 0; int ne = 0;
                const int e = ref(xinit, 4, bold - L, bold + L, dens_cb, &d, lcg_rand, &g1, &x, &ne);
                struct { int e, ne; double x; long calls; } r{e, ne, x, g1.calls};
                if (write(fd[1], &r, sizeof r) != (ssize_t)sizeof r) _exit(4);
                _exit(0);
            }
            close(fd[1]);
            struct { int e, ne; double x; long calls; } r{0, 0, 0, 0};
            const ssize_t got = read(fd[0], &r, sizeof r);
            close(fd[0]);
            int st = 0;
            waitpid(pid, &st, 0);
            if (got != (ssize_t)sizeof r) { e1 = 3000; } else { e1 = r.e; xr = r.x; nev = r.ne; g1.calls = r.calls; }
        }
        // restatement
        Lcg g2{seed, 0};
        hb::ArmsEnvelope env;
        double xm = 0;
        auto ur = [&]() { return (lcg_rand(&g2) + 0.5) / 2147483648.0; };
        const int e2 = hb::bw_sample_beta(m, Ck, sumSG, bold, ur, &xm, env);
        n++;
        if (e1 != 0 && e2 != 0 && (e1 == e2 || e1 == 3000)) { fail_both++; continue; }
        evals += nev;
        if (e1 != e2 || xr != xm || g1.calls != g2.calls || nev != env.neval) {
            if (bad < 5) printf("MISMATCH t=%d err %d/%d x %.17g/%.17g rand %ld/%ld neval %d/%d\n", t, e1, e2, xr, xm, g1.calls, g2.calls, nev, env.neval);
            bad++;
        }
    }
    printf("arms_check: %d cases, %d mismatches, %d rejected by both as not log-concave, mean evaluations %.2f\n", n, bad, fail_both, (double)evals / n);
    return bad ? 1 : 0;
}
