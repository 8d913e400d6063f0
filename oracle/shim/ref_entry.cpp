/*
 * ref_entry.cpp -- plain-C entry points around the reference's RcppExport functions as compiled against the header
 * shim (oracle/shim/RcppArmadillo.h): builds the SEXP arguments from plain arrays, calls the reference's own function,
 * copies the result out.  TEST INFRASTRUCTURE ONLY (oracle/_ref/libscde_ref.so); see oracle/Makefile target `ref`.
 * Argument layouts match oracle/scde_oracle.c's orc_* functions so the two can be called side by side.
 */
#include "jpmatLogBoot.h"
#include "matSlideMult.h"

namespace {

struct ArenaGuard {
    ~ArenaGuard() { ShimArena::get().clear(); }
};
SEXP mk_int(int v) {
    SEXP s = ShimArena::get().make(SHIM_INTSXP);
    s->ival.assign(1, v);
    return s;
}
SEXP mk_ivec(const int *p, size_t n) {
    SEXP s = ShimArena::get().make(SHIM_INTSXP);
    s->ival.assign(p, p + n);
    return s;
}
SEXP mk_imat(const int *p, int r, int c) {
    SEXP s = mk_ivec(p, (size_t)r * c);
    s->nrow = r;
    s->ncol = c;
    return s;
}
SEXP mk_dvec(const double *p, size_t n) {
    SEXP s = ShimArena::get().make(SHIM_REALSXP);
    s->dval.assign(p, p + n);
    return s;
}
SEXP mk_dmat(const double *p, int r, int c) {
    SEXP s = mk_dvec(p, (size_t)r * c);
    s->nrow = r;
    s->ncol = c;
    return s;
}
SEXP mk_list() { return ShimArena::get().make(SHIM_VECSXP); }
SEXP ucl_list(const int *flat, const int *off, int ncells) {
    SEXP l = mk_list();
    for (int c = 0; c < ncells; ++c) l->list.push_back(mk_ivec(flat + off[c], (size_t)(off[c + 1] - off[c])));
    return l;
}
SEXP named(SEXP l, const char *name) {
    for (size_t i = 0; i < l->names.size(); ++i)
        if (l->names[i] == name) return l->list[i];
    return nullptr;
}
/* result of logBoot(Batch)Posterior: REALSXP (jp) or a named list jp[, modes][, post] */
int unpack(SEXP r, int ngenes, int K, int ncells, double *jp, double *modes, double *post) {
    SEXP j = r->type == SHIM_VECSXP ? named(r, "jp") : r;
    if (!j || (int)j->dval.size() != ngenes * K) return -1;
    std::memcpy(jp, j->dval.data(), sizeof(double) * j->dval.size());
    if (r->type != SHIM_VECSXP) return 0;
    SEXP m = named(r, "modes");
    if (m && modes) std::memcpy(modes, m->dval.data(), sizeof(double) * m->dval.size());
    SEXP p = named(r, "post");
    if (p && post)
        for (int c = 0; c < ncells; ++c) /* each element G x K column-major; stored here as [cell][K][G] like the oracle */
            std::memcpy(post + (size_t)c * ngenes * K, p->list[(size_t)c]->dval.data(), sizeof(double) * (size_t)ngenes * K);
    return 0;
}

}  // namespace

extern "C" {

/* .Call("logBootPosterior", mm, ucl, uci, marginals, n.randomizations, Seed, postflag, ltheta, sqlogit, ensemble)
 * -- R/functions.R:637.  The draws come from srand(seed)/rand() inside the reference. */
int ref_log_boot_posterior(const double *models, int ncells, const int *ucl_flat, const int *ucl_off, const int *uci,
                           int ngenes, const double *mag, int K, int nboot, int seed, int returnpost, int localtheta,
                           int sqlogit, int ensemble, double *jp, double *modes, double *post) {
    ArenaGuard guard;
    SEXP r = logBootPosterior(mk_dmat(models, ncells, 12), ucl_list(ucl_flat, ucl_off, ncells), mk_imat(uci, ngenes, ncells),
                              mk_dvec(mag, (size_t)K), mk_int(nboot), mk_int(seed), mk_int(returnpost), mk_int(localtheta),
                              mk_int(sqlogit), mk_int(ensemble));
    return unpack(r, ngenes, K, ncells, jp, modes, post);
}

/* .Call("logBootBatchPosterior", mm, ucl, uci, marginals, batchil, composition, n.randomizations, Seed, postflag, ltheta,
 * sqlogit) -- R/functions.R:635 */
int ref_log_boot_batch_posterior(const double *models, int ncells, const int *ucl_flat, const int *ucl_off, const int *uci,
                                 int ngenes, const double *mag, int K, int nlevels, const int *pool_off,
                                 const int *pool_cells, const int *comp, int nboot, int seed, int returnpost, int localtheta,
                                 int sqlogit, double *jp, double *modes, double *post) {
    ArenaGuard guard;
    SEXP bl = mk_list();
    for (int k = 0; k < nlevels; ++k) bl->list.push_back(mk_ivec(pool_cells + pool_off[k], (size_t)(pool_off[k + 1] - pool_off[k])));
    SEXP r = logBootBatchPosterior(mk_dmat(models, ncells, 12), ucl_list(ucl_flat, ucl_off, ncells),
                                   mk_imat(uci, ngenes, ncells), mk_dvec(mag, (size_t)K), bl, mk_ivec(comp, (size_t)nlevels),
                                   mk_int(nboot), mk_int(seed), mk_int(returnpost), mk_int(localtheta), mk_int(sqlogit));
    return unpack(r, ngenes, K, ncells, jp, modes, post);
}

/* matl: nmat matrices nrows x ncols column-major, back to back (R/functions.R:3535) */
int ref_jpmat_log_boot(const double *matl, int nmat, int nrows, int ncols, int nboot, int seed, double *jp) {
    ArenaGuard guard;
    SEXP l = mk_list();
    for (int m = 0; m < nmat; ++m) l->list.push_back(mk_dmat(matl + (size_t)m * nrows * ncols, nrows, ncols));
    SEXP r = jpmatLogBoot(l, mk_int(nboot), mk_int(seed));
    std::memcpy(jp, r->dval.data(), sizeof(double) * (size_t)nrows * ncols);
    return 0;
}

/* pools back to back; pool k = matrices pool_off[k] .. pool_off[k+1] (R/functions.R:3541) */
int ref_jpmat_log_batch_boot(const double *matl, int nlevels, const int *pool_off, const int *comp, int nrows, int ncols,
                             int nboot, int seed, double *jp) {
    ArenaGuard guard;
    SEXP ll = mk_list();
    for (int k = 0; k < nlevels; ++k) {
        SEXP l = mk_list();
        for (int m = pool_off[k]; m < pool_off[k + 1]; ++m) l->list.push_back(mk_dmat(matl + (size_t)m * nrows * ncols, nrows, ncols));
        ll->list.push_back(l);
    }
    SEXP r = jpmatLogBatchBoot(ll, mk_ivec(comp, (size_t)nlevels), mk_int(nboot), mk_int(seed));
    std::memcpy(jp, r->dval.data(), sizeof(double) * (size_t)nrows * ncols);
    return 0;
}

/* .Call("matSlideMult", m1, m2) -- R/functions.R:3545 */
int ref_mat_slide_mult(const double *m1, const double *m2, int nrows, int n, double *out) {
    ArenaGuard guard;
    SEXP r = matSlideMult(mk_dmat(m1, nrows, n), mk_dmat(m2, nrows, n));
    std::memcpy(out, r->dval.data(), sizeof(double) * (size_t)nrows * (2 * n - 1));
    return 0;
}

}  // extern "C"
