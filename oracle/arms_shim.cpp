// Test-infrastructure shim around the REFERENCE's own ARMS sampler (src/BayesW_arms.cpp), which
// oracle/build_ref.sh compiles from where it lies and links with --wrap=rand, so that the uniforms it consumes
// (u = (rand()+0.5)/(RAND_MAX+1), src/BayesW_arms.cpp:913-918) come from a caller-supplied source.
// Not product code; lives in oracle/_ref/libarms_ref.so together with the reference object code.
#include <stdlib.h>

// the reference's entry point (src/BayesW_arms.cpp:135)
int arms(double *xinit, int ninit, double *xl, double *xr, double (*myfunc)(double x, void *mydata), void *mydata, double *convex,
         int npoint, int dometrop, double *xprev, double *xsamp, int nsamp, double *qcent, double *xcent, int ncent, int *neval);

static int (*g_rand_fn)(void *) = nullptr;
static void *g_rand_data = nullptr;
static long g_rand_calls = 0;

extern "C" int __wrap_rand(void) {  // what the reference's rand() calls resolve to (ld --wrap=rand)
    g_rand_calls++;
    return g_rand_fn ? g_rand_fn(g_rand_data) : 0;
}

extern "C" __attribute__((visibility("default"))) long ho_ref_arms_rand_calls(void) { return g_rand_calls; }

// One ARMS draw with hydra's settings (ninit=4, npoint=100, nsamp=1, convex=1, dometrop=0, the four ignored centiles:
// src/BayesW.cpp:1336-1343). Returns the reference's error code.
extern "C" __attribute__((visibility("default"))) int ho_ref_arms(const double *xinit, int ninit, double xl, double xr,
                                                                  double (*dens)(double, void *), void *data, int (*rand_fn)(void *),
                                                                  void *rand_data, double *xsamp_out, int *neval_out) {
    double xi[16];
    for (int i = 0; i < ninit && i < 16; i++) xi[i] = xinit[i];
    double convex = 1.0, xprev = 0.0, xsamp[1] = {0.0}, xcent[10], qcent[10] = {5., 30., 70., 95.};
    int neval = 0;
    g_rand_fn = rand_fn;
    g_rand_data = rand_data;
    const int err = arms(xi, ninit, &xl, &xr, dens, data, &convex, 100, 0, &xprev, xsamp, 1, qcent, xcent, 4, &neval);
    g_rand_fn = nullptr;
    *xsamp_out = xsamp[0];
    if (neval_out) *neval_out = neval;
    return err;
}
