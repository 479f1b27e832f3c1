// pn_fenrir_update.cuh -- Fenrir data update on one trajectory's posterior, executed by a single thread.
//
// Reference behaviour replaced: measure_and_update! (src/data_likelihoods/fenrir.jl:127-137) = make_obscov_sqrt
// (src/callbacks/dataupdate.jl:86-87) + update!(x, x_pred, msmnt, H; R) (src/filtering/update.jl:65-139).  The reference
// forms K, applies the Joseph form R (I - K H)' and re-triangularises the stack with (K Rn')' (update.jl:132-136).  Here the
// same posterior comes out of ONE orthogonal transformation of the pre-array
//
//        [ Rn   0 ]            [ S.R   Kb ]        S   = S.R' S.R            (innovation covariance H Sigma H' + Rn'Rn)
//    Q   [        ]     =      [          ]        Kb  = S.R^-T H Sigma      (so Sigma H' S^-1 z = Kb' S.R^-T z)
//        [ R H'  R ]           [  0    R+ ]        R+' R+ = Sigma - Sigma H' S^-1 H Sigma   (the updated covariance)
//
// i.e. n_obs Householder reflections over D + n_obs rows instead of a (D + n_obs) x D QR plus three D x D products
// (about a third of the arithmetic; identical in exact arithmetic, orthogonal transformations only).  R+ is a dense right
// factor -- the sweep kernels accept that (the factor of the last point is dense as well).
//
// Plain C++ on purpose (no CUDA builtins): tests/test_host.py compiles this header with g++ and checks it against the
// reference formulas.  PN_FEN_FN carries the function qualifiers of the device build.
#ifndef PN_FENRIR_UPDATE_CUH
#define PN_FENRIR_UPDATE_CUH

#ifndef PN_FEN_FN
#define PN_FEN_FN __device__ __noinline__
#endif
#ifndef PN_FEN_INF
#define PN_FEN_INF PN_INF
#endif

// Posterior mean Ms[D] and right factor Sm (D rows, leading dimension LD) are updated in place on the observation u[o] with
// observation matrix H (o x D row-major) and noise right factor Rn (o x o row-major, upper triangular).  Returns the
// log-likelihood of the observation; +-Inf without touching the state if Sigma == 0 (update.jl:82-89).  kron_logdet: the
// reference's value on the IsometricKronecker layout (log det S / o, see PnSmoothParams::kron_logdet).
template <int D, int LD, int OMAX>
PN_FEN_FN double pn_fenrir_update_serial(const double* H, const double* Rn, const double* u, int o, double* Sm, double* Ms,
                                         bool kron_logdet = false) {
    double Top[OMAX][OMAX + D]; /* rows 0..o-1 of the pre-array: [Rn | 0] -> [S.R | Kb] */
    double Cb[D][OMAX];         /* left block of the bottom rows: R H' -> 0               */
    double z[OMAX], w[OMAX];
    double nrm = 0.0;
    for (int r = 0; r < D; r++)
        for (int c = 0; c < D; c++) nrm += fabs(Sm[r * LD + c]);
    for (int a = 0; a < o; a++) {
        double sacc = 0.0;
        for (int c = 0; c < D; c++) sacc += H[a * D + c] * Ms[c];
        z[a] = sacc - u[a];
    }
    if (nrm == 0.0) {
        double nz = 0.0;
        for (int a = 0; a < o; a++) nz += fabs(z[a]);
        return (nz == 0.0) ? PN_FEN_INF : -PN_FEN_INF;
    }
    for (int a = 0; a < o; a++) {
        for (int b = 0; b < o; b++) Top[a][b] = Rn[a * o + b];
        for (int c = 0; c < D; c++) Top[a][o + c] = 0.0;
    }
    for (int r = 0; r < D; r++)
        for (int a = 0; a < o; a++) {
            double sacc = 0.0;
            for (int c = 0; c < D; c++) sacc += Sm[r * LD + c] * H[a * D + c];
            Cb[r][a] = sacc;
        }
    /* o Householder reflections: column k, rows k..o-1 of Top and all D bottom rows */
    for (int k = 0; k < o; k++) {
        double total = 0.0;
        for (int a = k; a < o; a++) total += Top[a][k] * Top[a][k];
        for (int r = 0; r < D; r++) total += Cb[r][k] * Cb[r][k];
        if (!(total > 0.0)) continue;
        const double alpha = Top[k][k], nr = sqrt(total), beta = -copysign(nr, alpha), vk = alpha - beta;
        const double gam = 1.0 / (total + fabs(alpha) * nr);
        for (int j = k + 1; j < o + D; j++) { /* columns right of k: first the remaining Top/Cb columns, then the R block */
            double sj = vk * Top[k][j];
            for (int a = k + 1; a < o; a++) sj += Top[a][k] * Top[a][j];
            if (j < o) {
                for (int r = 0; r < D; r++) sj += Cb[r][k] * Cb[r][j];
            } else {
                for (int r = 0; r < D; r++) sj += Cb[r][k] * Sm[r * LD + (j - o)];
            }
            sj *= gam;
            Top[k][j] -= sj * vk;
            for (int a = k + 1; a < o; a++) Top[a][j] -= sj * Top[a][k];
            if (j < o) {
                for (int r = 0; r < D; r++) Cb[r][j] -= sj * Cb[r][k];
            } else {
                for (int r = 0; r < D; r++) Sm[r * LD + (j - o)] -= sj * Cb[r][k];
            }
        }
        Top[k][k] = beta;
        for (int a = k + 1; a < o; a++) Top[a][k] = 0.0;
        for (int r = 0; r < D; r++) Cb[r][k] = 0.0; /* (not read again; kept for clarity of the invariant) */
    }
    /* w = S.R^-T z (forward substitution), log-likelihood, mean */
    double quad = 0.0, logdet = 0.0;
    for (int a = 0; a < o; a++) {
        double sacc = z[a];
        for (int b = 0; b < a; b++) sacc -= Top[b][a] * w[b];
        w[a] = sacc / Top[a][a];
        quad += w[a] * w[a];
        logdet += log(fabs(Top[a][a]));
    }
    for (int c = 0; c < D; c++) {
        double sacc = Ms[c];
        for (int a = 0; a < o; a++) sacc -= Top[a][o + c] * w[a];
        Ms[c] = sacc;
    }
    /* IsometricKronecker layout: logdet of the 1 x 1 factor of S = s I, not multiplied by d (update.jl:140-148, 182-194) */
    if (kron_logdet) logdet /= (double)o;
    return -0.5 * quad - 0.5 * (double)o * 1.8378770664093453 - logdet;
}

#endif /* PN_FENRIR_UPDATE_CUH */
