// pn_ioup.cuh -- per-step transition quantities of the integrated Ornstein-Uhlenbeck prior with a scalar rate
// parameter r (src/priors/ioup.jl:66-79: dY^(q) = r Y^(q) dt + dW).
//
// Reference behaviour replaced: make_transition_matrices!(cache, ::IOUP, dt) (src/priors/ioup.jl:115-131), which calls
// FiniteHorizonGramians.exp_and_gram_chol! on (F, L, dt) and then forms A = P Ah PI, Q = P Qh P.  In the preconditioned
// coordinates x~ = P x (src/preconditioning.jl:30-58) the process only depends on z = r h:
//     A~[a][b] = binomial(q-a, q-b)               (b < q, as for the IWP, src/priors/iwp.jl:99)
//     A~[a][q] = m! phi_m(z),  m = q - a,          phi_m(z) = sum_k z^k / (k+m)!
//     Q~[a][b] = m_a! m_b! int_0^1 tau^(m_a+m_b) phi_(m_a)(z tau) phi_(m_b)(z tau) dtau     (Hilbert matrix at z = 0)
// i.e. in filter coordinates Ah[a][q] = h^m phi_m(z) and Qh.R = chol(Q~).U P^-1.  The integral is a 16-point
// Gauss-Legendre rule (nodes/weights host-computed, exact to rounding for |z| <~ 8), distributed over the G lanes of
// the trajectory's group and reduced with group-masked butterflies, so all lanes of the group hold bitwise identical
// results (the groups of a warp may be in different retry iterations when this runs).
//
// NVRTC-compatible: no standard headers.
#ifndef PN_IOUP_CUH
#define PN_IOUP_CUH
#include "pn_params.h"
#include "pn_small.cuh"

template <int Q>
PN_HD void pn_phi_all(double x, double (&ph)[Q + 1]) {
    double ifact[Q + 1]; /* 1/m! */
    ifact[0] = 1.0;
#pragma unroll
    for (int m = 1; m <= Q; m++) ifact[m] = ifact[m - 1] / (double)m;
    if (fabs(x) < 1.0) {
        double s = 0.0, term = ifact[Q];
#pragma unroll
        for (int k = 0; k < 20; k++) {
            s += term;
            term *= x / (double)(k + 1 + Q);
        }
        ph[Q] = s;
#pragma unroll
        for (int m = Q; m >= 1; m--) ph[m - 1] = fma(x, ph[m], ifact[m - 1]);
    } else {
        const double ix = 1.0 / x;
        ph[0] = exp(x);
#pragma unroll
        for (int m = 1; m <= Q; m++) ph[m] = (ph[m - 1] - ifact[m - 1]) * ix;
    }
}

// tq[j] = Ah[j][q] (one component) = h^(q-j) phi_(q-j)(z);  Uh = chol(Q~).U (upper, row-major n x n, lower part zero).
// Returns false if Q~ is not numerically positive definite.
template <int Q, int G>
PN_HD bool pn_ioup_step(const double rate, const double* glx, const double* glw, double h, int gl, double (&tq)[Q + 1],
                        double (&Uh)[Q + 1][Q + 1]) {
    constexpr int n = Q + 1;
    const double z = rate * h;
    double fact[n];
    fact[0] = 1.0;
#pragma unroll
    for (int m = 1; m < n; m++) fact[m] = fact[m - 1] * (double)m;
    const unsigned gmask = (G >= 32) ? 0xffffffffu : (((1u << G) - 1u) << ((threadIdx.x & 31) - gl));
    double Qt[n][n];
#pragma unroll
    for (int a = 0; a < n; a++)
#pragma unroll
        for (int b = 0; b < n; b++) Qt[a][b] = 0.0;
    for (int i = gl; i < 16; i += G) {
        const double tau = glx[i], w = glw[i];
        double ph[n], g[n];
        pn_phi_all<Q>(z * tau, ph);
        double tp = 1.0; /* tau^m */
#pragma unroll
        for (int m = 0; m < n; m++) {
            g[Q - m] = fact[m] * tp * ph[m];
            tp *= tau;
        }
#pragma unroll
        for (int a = 0; a < n; a++)
#pragma unroll
            for (int b = 0; b < n; b++)
                if (b >= a) Qt[a][b] += w * g[a] * g[b];
    }
#pragma unroll
    for (int a = 0; a < n; a++)
#pragma unroll
        for (int b = 0; b < n; b++)
            if (b >= a) { /* group-masked butterflies: the step kernel calls this from its divergent retry phase */
                double v = Qt[a][b];
#pragma unroll
                for (int o = G >> 1; o >= 1; o >>= 1) v += __shfl_xor_sync(gmask, v, o, 32);
                Qt[a][b] = v;
            }
    double inv[n];
    const bool ok = pn_chol<n>(Qt, inv);
#pragma unroll
    for (int a = 0; a < n; a++)
#pragma unroll
        for (int b = 0; b < n; b++) Uh[a][b] = (b >= a) ? Qt[a][b] : 0.0;
    double ph[n];
    pn_phi_all<Q>(z, ph);
    double hpw = 1.0; /* h^m */
#pragma unroll
    for (int m = 0; m < n; m++) {
        tq[Q - m] = hpw * ph[m];
        hpw *= h;
    }
    return ok;
}

#endif /* PN_IOUP_CUH */
