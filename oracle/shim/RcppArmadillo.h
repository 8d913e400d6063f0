/*
 * RcppArmadillo.h -- a header SHIM, test infrastructure only (never linked into libscde_b200).
 *
 * Purpose: compile the reference's own, UNMODIFIED translation units
 *     /root/reference/src/jpmatLogBoot.cpp   (jpmatLogBoot, jpmatLogBatchBoot, logBootPosterior, logBootBatchPosterior)
 *     /root/reference/src/matSlideMult.cpp   (matSlideMult)
 * in an image that has neither R, Rcpp, RcppArmadillo nor Armadillo, so that the oracle's
 * restatement of their loop nests can be checked against "the reference compiled here" (oracle/_ref/libscde_ref.so,
 * recipe: oracle/Makefile target `ref`).  The reference sources are compiled where they lie; nothing is copied.
 *
 * What is provided is exactly the part of the three libraries those two files use, with the semantics of the real ones:
 *   - R API:   SEXP (INTSXP / REALSXP / VECSXP with dim and names), VECTOR_ELT, LENGTH, R_CheckUserInterrupt (no-op),
 *              Rf_dnbinom / Rf_dpois = the oracle's restatement of R nmath (oracle/scde_oracle.c; pinned by mpmath in
 *              tests/test_oracle.py -- R itself is absent, so nmath stays "restated")
 *   - Rcpp:    as<int|IntegerVector|IntegerMatrix|NumericMatrix|arma::mat|arma::colvec>, wrap, List, Named, List::create
 *   - Armadillo: mat / vec / colvec / rowvec (column-major, operator(), [], n_elem, n_rows, n_cols, zeros, t, col, cols,
 *              each_col, each_row, max(index)), element-wise + - * / % with scalars and matrices, exp, log, exp10, pow,
 *              max, sum.  Operations are evaluated eagerly, element by element, with libm -- element-wise results do
 *              not depend on Armadillo's expression templates.  The REDUCTIONS follow Armadillo's own summation order,
 *              which is what decides the last bit:
 *                sum(vector), sum(M, 0)    arrayops::accumulate: two interleaved partial sums, acc1 + acc2
 *                sum(M, 1), sum(A % B, 1)  column after column added into the output (op_sum, dim = 1)
 */
#ifndef SCDE_ORACLE_SHIM_RCPPARMADILLO_H
#define SCDE_ORACLE_SHIM_RCPPARMADILLO_H

#include <cmath>
#include <cstddef>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <string>
#include <vector>
#include <algorithm>
#include <iostream>
#include <iterator>
#include <stdexcept>

/* ---- R API ------------------------------------------------------------------------------------------------------ */
enum { SHIM_INTSXP = 13, SHIM_REALSXP = 14, SHIM_VECSXP = 19 };
struct ShimSexp {
    int type = SHIM_REALSXP;
    std::vector<int> ival;
    std::vector<double> dval;
    std::vector<ShimSexp *> list;
    std::vector<std::string> names;
    int nrow = -1, ncol = -1; /* dim attribute, -1 = none */
};
typedef ShimSexp *SEXP;

/* every SEXP made during one call lives in an arena the entry wrapper frees (there is no garbage collector) */
struct ShimArena {
    std::vector<ShimSexp *> all;
    ShimSexp *make(int type) {
        ShimSexp *s = new ShimSexp();
        s->type = type;
        all.push_back(s);
        return s;
    }
    void clear() {
        for (ShimSexp *s : all) delete s;
        all.clear();
    }
    static ShimArena &get() {
        static thread_local ShimArena a;
        return a;
    }
};

inline SEXP VECTOR_ELT(SEXP x, std::ptrdiff_t i) { return x->list[(size_t)i]; }
inline int LENGTH(SEXP x) {
    return x->type == SHIM_VECSXP ? (int)x->list.size() : (x->type == SHIM_INTSXP ? (int)x->ival.size() : (int)x->dval.size());
}
inline void R_CheckUserInterrupt() {}
extern "C" double orc_dnbinom_log(double x, double size, double prob); /* oracle/scde_oracle.c */
extern "C" double orc_dpois_log(double x, double lambda);
inline double Rf_dnbinom(double x, double size, double prob, int give_log) {
    const double v = orc_dnbinom_log(x, size, prob);
    return give_log ? v : std::exp(v);
}
inline double Rf_dpois(double x, double lambda, int give_log) {
    const double v = orc_dpois_log(x, lambda);
    return give_log ? v : std::exp(v);
}
#define RcppExport extern "C"

/* ---- Armadillo -------------------------------------------------------------------------------------------------- */
namespace arma {
typedef unsigned long long uword;

inline double shim_accumulate(const double *src, uword n) { /* arrayops::accumulate */
    double acc1 = 0.0, acc2 = 0.0;
    uword j;
    for (j = 1; j < n; j += 2) {
        acc1 += *src++;
        acc2 += *src++;
    }
    if ((j - 1) < n) acc1 += *src;
    return acc1 + acc2;
}

class Mat;
struct ColRef;      /* M.col(j): one column, read/write */
struct ColsRef;     /* M.cols(a, b): columns a..b, read only */
struct EachCol;
struct EachRow;

class Mat {
  public:
    uword n_rows = 0, n_cols = 0, n_elem = 0;

    Mat() {}
    Mat(uword r, uword c) { init(r, c); }
    /* auxiliary-memory constructor: copy_aux_mem = false uses the caller's memory in place */
    Mat(double *aux, uword r, uword c, bool copy_aux_mem = true, bool /*strict*/ = false) {
        n_rows = r;
        n_cols = c;
        n_elem = r * c;
        if (copy_aux_mem) {
            own.assign(aux, aux + n_elem);
            mem = own.data();
        } else {
            mem = aux;
        }
    }
    Mat(const Mat &o) { copy_from(o); }
    Mat &operator=(const Mat &o) {
        if (this != &o) copy_from(o);
        return *this;
    }
    Mat(const ColRef &c);
    Mat &operator=(const ColRef &c);

    double *memptr() { return mem; }
    const double *memptr() const { return mem; }
    double *colptr(uword c) { return mem + c * n_rows; }
    const double *colptr(uword c) const { return mem + c * n_rows; }
    double &operator()(uword r, uword c) { return mem[r + c * n_rows]; }
    double operator()(uword r, uword c) const { return mem[r + c * n_rows]; }
    double &operator[](uword i) { return mem[i]; }
    double operator[](uword i) const { return mem[i]; }
    double &at(uword r, uword c) { return mem[r + c * n_rows]; }
    double *begin() { return mem; }
    double *end() { return mem + n_elem; }
    const double *begin() const { return mem; }
    const double *end() const { return mem + n_elem; }

    void zeros() { std::fill(mem, mem + n_elem, 0.0); }
    void set_size(uword r, uword c) { init(r, c); }
    Mat t() const {
        Mat o(n_cols, n_rows);
        for (uword c = 0; c < n_cols; ++c)
            for (uword r = 0; r < n_rows; ++r) o.mem[c + r * n_cols] = mem[r + c * n_rows];
        return o;
    }
    double max(uword &index) const { /* first maximum, as op_max::direct_max with an index */
        double best = -std::numeric_limits<double>::infinity();
        uword bi = 0;
        for (uword i = 0; i < n_elem; ++i)
            if (mem[i] > best) {
                best = mem[i];
                bi = i;
            }
        index = bi;
        return best;
    }
    inline ColRef col(uword j);
    inline const ColRef col(uword j) const;
    inline ColsRef cols(uword a, uword b) const;
    inline EachCol each_col();
    inline EachRow each_row();

    Mat &operator+=(const Mat &o) {
        for (uword i = 0; i < n_elem; ++i) mem[i] += o.mem[i];
        return *this;
    }
    Mat &operator%=(const Mat &o) {
        for (uword i = 0; i < n_elem; ++i) mem[i] *= o.mem[i];
        return *this;
    }
    Mat &operator/=(const Mat &o) {
        for (uword i = 0; i < n_elem; ++i) mem[i] /= o.mem[i];
        return *this;
    }
    Mat &operator+=(double k) {
        for (uword i = 0; i < n_elem; ++i) mem[i] += k;
        return *this;
    }
    Mat &operator*=(double k) {
        for (uword i = 0; i < n_elem; ++i) mem[i] *= k;
        return *this;
    }
    Mat &operator/=(double k) {
        for (uword i = 0; i < n_elem; ++i) mem[i] /= k;
        return *this;
    }

  protected:
    std::vector<double> own;
    double *mem = nullptr;
    void init(uword r, uword c) {
        n_rows = r;
        n_cols = c;
        n_elem = r * c;
        own.assign((size_t)n_elem, 0.0); /* Armadillo leaves new memory uninitialised; every use here writes before it reads */
        mem = own.data();
    }
    void copy_from(const Mat &o) {
        n_rows = o.n_rows;
        n_cols = o.n_cols;
        n_elem = o.n_elem;
        own.assign(o.mem, o.mem + o.n_elem);
        mem = own.data();
    }
};
typedef Mat mat;

/* Col / Row: a Mat with one column / one row; `vec(n)` and `colvec(n)` size a column */
class Col : public Mat {
  public:
    Col() : Mat() {}
    explicit Col(uword n) : Mat(n, 1) {}
    Col(const Mat &m) : Mat(m) {}
    Col(const ColRef &c) : Mat(c) {}
    Col &operator=(const Mat &m) {
        Mat::operator=(m);
        return *this;
    }
};
class Row : public Mat {
  public:
    Row() : Mat() {}
    explicit Row(uword n) : Mat(1, n) {}
    Row(const Mat &m) : Mat(m) {}
    Row &operator=(const Mat &m) {
        Mat::operator=(m);
        return *this;
    }
};
typedef Col vec;
typedef Col colvec;
typedef Row rowvec;

struct ColRef {
    Mat *m;
    uword j;
    double *ptr() const { return m->colptr(j); }
    uword n() const { return m->n_rows; }
    const ColRef &operator=(const Mat &v) const {
        std::memcpy(ptr(), v.memptr(), sizeof(double) * (size_t)n());
        return *this;
    }
    const ColRef &operator=(const ColRef &v) const {
        std::memmove(ptr(), v.ptr(), sizeof(double) * (size_t)n());
        return *this;
    }
    const ColRef &operator+=(const ColRef &v) const {
        double *d = ptr();
        const double *s = v.ptr();
        for (uword i = 0; i < n(); ++i) d[i] += s[i];
        return *this;
    }
};
inline Mat::Mat(const ColRef &c) {
    init(c.n(), 1);
    std::memcpy(mem, c.ptr(), sizeof(double) * (size_t)n_elem);
}
inline Mat &Mat::operator=(const ColRef &c) {
    init(c.n(), 1);
    std::memcpy(mem, c.ptr(), sizeof(double) * (size_t)n_elem);
    return *this;
}
inline ColRef Mat::col(uword j) { return ColRef{this, j}; }
inline const ColRef Mat::col(uword j) const { return ColRef{const_cast<Mat *>(this), j}; }

struct ColsRef {
    const Mat *m;
    uword a, b; /* inclusive */
    uword n_rows() const { return m->n_rows; }
    uword n_cols() const { return b - a + 1; }
    double at(uword r, uword c) const { return (*m)(r, a + c); }
};
inline ColsRef Mat::cols(uword a, uword b) const { return ColsRef{this, a, b}; }
/* A.cols(..) % B.cols(..): kept lazy, as Armadillo's eGlue is; only sum(.., 1) consumes it */
struct ColsSchur {
    ColsRef x, y;
};
inline ColsSchur operator%(const ColsRef &x, const ColsRef &y) { return ColsSchur{x, y}; }

struct EachCol {
    Mat *m;
    void operator-=(const Mat &v) const {
        for (uword c = 0; c < m->n_cols; ++c)
            for (uword r = 0; r < m->n_rows; ++r) (*m)(r, c) -= v[r];
    }
    void operator/=(const Mat &v) const {
        for (uword c = 0; c < m->n_cols; ++c)
            for (uword r = 0; r < m->n_rows; ++r) (*m)(r, c) /= v[r];
    }
};
struct EachRow {
    Mat *m;
    void operator-=(const Mat &v) const {
        for (uword c = 0; c < m->n_cols; ++c)
            for (uword r = 0; r < m->n_rows; ++r) (*m)(r, c) -= v[c];
    }
    void operator/=(const Mat &v) const {
        for (uword c = 0; c < m->n_cols; ++c)
            for (uword r = 0; r < m->n_rows; ++r) (*m)(r, c) /= v[c];
    }
};
inline EachCol Mat::each_col() { return EachCol{this}; }
inline EachRow Mat::each_row() { return EachRow{this}; }

/* element-wise helpers */
template <class F>
inline Mat shim_map(const Mat &x, F f) {
    Mat o(x.n_rows, x.n_cols);
    for (uword i = 0; i < x.n_elem; ++i) o[i] = f(x[i]);
    return o;
}
template <class F>
inline Mat shim_zip(const Mat &x, const Mat &y, F f) {
    Mat o(x.n_rows, x.n_cols);
    for (uword i = 0; i < x.n_elem; ++i) o[i] = f(x[i], y[i]);
    return o;
}
inline Mat operator*(const Mat &x, double k) { return shim_map(x, [k](double v) { return v * k; }); }
inline Mat operator*(double k, const Mat &x) { return shim_map(x, [k](double v) { return v * k; }); } /* eop_scalar_times */
inline Mat operator+(const Mat &x, double k) { return shim_map(x, [k](double v) { return v + k; }); }
inline Mat operator+(double k, const Mat &x) { return shim_map(x, [k](double v) { return v + k; }); } /* eop_scalar_plus */
inline Mat operator-(const Mat &x, double k) { return shim_map(x, [k](double v) { return v - k; }); }
inline Mat operator-(double k, const Mat &x) { return shim_map(x, [k](double v) { return k - v; }); }
inline Mat operator/(double k, const Mat &x) { return shim_map(x, [k](double v) { return k / v; }); }
inline Mat operator/(const Mat &x, double k) { return shim_map(x, [k](double v) { return v / k; }); }
inline Mat operator+(const Mat &x, const Mat &y) { return shim_zip(x, y, [](double a, double b) { return a + b; }); }
inline Mat operator-(const Mat &x, const Mat &y) { return shim_zip(x, y, [](double a, double b) { return a - b; }); }
inline Mat operator%(const Mat &x, const Mat &y) { return shim_zip(x, y, [](double a, double b) { return a * b; }); }
inline Mat exp(const Mat &x) { return shim_map(x, [](double v) { return std::exp(v); }); }
inline Mat log(const Mat &x) { return shim_map(x, [](double v) { return std::log(v); }); }
inline Mat exp10(const Mat &x) { return shim_map(x, [](double v) { return std::pow(10.0, v); }); } /* eop_aux::exp10 */
inline Mat pow(const Mat &x, double k) { return shim_map(x, [k](double v) { return std::pow(v, k); }); }

/* reductions */
inline double max(const Col &v) { /* op_max::direct_max: starts from -inf, strict >, so NaN entries are skipped */
    double best = -std::numeric_limits<double>::infinity();
    for (uword i = 0; i < v.n_elem; ++i)
        if (v[i] > best) best = v[i];
    return best;
}
inline double sum(const Col &v) { return shim_accumulate(v.memptr(), v.n_elem); }
inline Mat max(const Mat &x, int dim) {
    if (dim == 0) {
        Mat o(1, x.n_cols);
        for (uword c = 0; c < x.n_cols; ++c) {
            const double *p = x.colptr(c);
            double best = -std::numeric_limits<double>::infinity();
            for (uword r = 0; r < x.n_rows; ++r)
                if (p[r] > best) best = p[r];
            o[c] = best;
        }
        return o;
    }
    Mat o(x.n_rows, 1);
    for (uword r = 0; r < x.n_rows; ++r) o[r] = x(r, 0);
    for (uword c = 1; c < x.n_cols; ++c)
        for (uword r = 0; r < x.n_rows; ++r)
            if (x(r, c) > o[r]) o[r] = x(r, c);
    return o;
}
inline Mat sum(const Mat &x, int dim) {
    if (dim == 0) { /* one accumulate() per column */
        Mat o(1, x.n_cols);
        for (uword c = 0; c < x.n_cols; ++c) o[c] = shim_accumulate(x.colptr(c), x.n_rows);
        return o;
    }
    Mat o(x.n_rows, 1); /* out.zeros(); out += column, column after column */
    o.zeros();
    for (uword c = 0; c < x.n_cols; ++c) {
        const double *p = x.colptr(c);
        for (uword r = 0; r < x.n_rows; ++r) o[r] += p[r];
    }
    return o;
}
inline Mat sum(const ColsSchur &e, int dim) { /* op_sum::apply_noalias_proxy, dim = 1: out[row] += P.at(row, col) */
    (void)dim;
    const uword nr = e.x.n_rows(), nc = e.x.n_cols();
    Mat o(nr, 1);
    o.zeros();
    for (uword c = 0; c < nc; ++c)
        for (uword r = 0; r < nr; ++r) o[r] += e.x.at(r, c) * e.y.at(r, c);
    return o;
}
} /* namespace arma */

/* ---- Rcpp ------------------------------------------------------------------------------------------------------- */
namespace Rcpp {

class IntegerVector {
  public:
    SEXP s;
    IntegerVector(SEXP x) : s(x) {}
    int size() const { return (int)s->ival.size(); }
    int operator[](std::ptrdiff_t i) const { return s->ival[(size_t)i]; }
    int *begin() { return s->ival.data(); }
    int *end() { return s->ival.data() + s->ival.size(); }
};
class IntegerMatrix {
  public:
    SEXP s;
    IntegerMatrix(SEXP x) : s(x) {}
    int nrow() const { return s->nrow; }
    int ncol() const { return s->ncol; }
    int operator()(std::ptrdiff_t i, std::ptrdiff_t j) const { return s->ival[(size_t)(i + (std::ptrdiff_t)s->nrow * j)]; }
    int *begin() { return s->ival.data(); }
};
class NumericVector {
  public:
    SEXP s;
    NumericVector(SEXP x) : s(x) {}
    int size() const { return (int)s->dval.size(); }
    double operator[](std::ptrdiff_t i) const { return s->dval[(size_t)i]; }
    double *begin() { return s->dval.data(); }
};
struct ShimStop : public std::runtime_error { /* Rcpp::stop: an R error; here a C++ exception the entry wrapper reports */
    explicit ShimStop(const std::string &m) : std::runtime_error(m) {}
};
inline void stop(const std::string &msg) { throw ShimStop(msg); }
class NumericMatrix {
  public:
    SEXP s;
    NumericMatrix(SEXP x) : s(x) {}
    NumericMatrix(int r, int c) {
        s = ShimArena::get().make(SHIM_REALSXP);
        s->nrow = r;
        s->ncol = c;
        s->dval.assign((size_t)r * c, 0.0);
    }
    int nrow() const { return s->nrow; }
    int ncol() const { return s->ncol; }
    int size() const { return (int)s->dval.size(); }
    double *begin() { return s->dval.data(); }
    double &operator()(std::ptrdiff_t i, std::ptrdiff_t j) { return s->dval[(size_t)(i + (std::ptrdiff_t)s->nrow * j)]; }
    operator SEXP() const { return s; }
};

template <class T>
struct AsImpl;
template <>
struct AsImpl<int> {
    static int get(SEXP x) { return x->type == SHIM_INTSXP ? x->ival[0] : (int)x->dval[0]; }
};
template <>
struct AsImpl<IntegerVector> {
    static IntegerVector get(SEXP x) { return IntegerVector(x); }
};
template <>
struct AsImpl<IntegerMatrix> {
    static IntegerMatrix get(SEXP x) { return IntegerMatrix(x); }
};
template <>
struct AsImpl<NumericMatrix> {
    static NumericMatrix get(SEXP x) { return NumericMatrix(x); }
};
template <>
struct AsImpl<NumericVector> {
    static NumericVector get(SEXP x) { return NumericVector(x); }
};
template <>
struct AsImpl<arma::mat> {
    static arma::mat get(SEXP x) { return arma::mat(x->dval.data(), (arma::uword)x->nrow, (arma::uword)x->ncol, true); }
};
template <>
struct AsImpl<arma::colvec> {
    static arma::colvec get(SEXP x) { return arma::colvec(arma::mat(x->dval.data(), (arma::uword)x->dval.size(), 1, true)); }
};
template <class T>
inline T as(SEXP x) {
    return AsImpl<T>::get(x);
}

inline SEXP wrap(SEXP x) { return x; }
inline SEXP wrap(const arma::Mat &m) {
    SEXP s = ShimArena::get().make(SHIM_REALSXP);
    s->nrow = (int)m.n_rows;
    s->ncol = (int)m.n_cols;
    s->dval.assign(m.memptr(), m.memptr() + m.n_elem);
    return s;
}
inline SEXP wrap(const NumericMatrix &m) { return m.s; }

struct NamedValue {
    std::string name;
    SEXP value;
};
struct Named {
    std::string name;
    explicit Named(const char *n) : name(n) {}
    NamedValue operator=(SEXP v) const { return NamedValue{name, v}; }
};

class List {
  public:
    SEXP s;
    struct Slot {
        SEXP *p;
        operator SEXP() const { return *p; }
        Slot &operator=(SEXP v) {
            *p = v;
            return *this;
        }
        Slot &operator=(const arma::Mat &m) {
            *p = wrap(m);
            return *this;
        }
        Slot &operator=(const NumericMatrix &m) {
            *p = m.s;
            return *this;
        }
    };
    List(SEXP x) : s(x) {}
    explicit List(int n) {
        s = ShimArena::get().make(SHIM_VECSXP);
        s->list.assign((size_t)n, nullptr);
    }
    int size() const { return (int)s->list.size(); }
    Slot operator[](std::ptrdiff_t i) { return Slot{&s->list[(size_t)i]}; }
    operator SEXP() const { return s; }
    static SEXP make(std::initializer_list<NamedValue> nv) {
        SEXP s = ShimArena::get().make(SHIM_VECSXP);
        for (const NamedValue &x : nv) {
            s->list.push_back(x.value);
            s->names.push_back(x.name);
        }
        return s;
    }
    static SEXP create(const NamedValue &a) { return make({a}); }
    static SEXP create(const NamedValue &a, const NamedValue &b) { return make({a, b}); }
    static SEXP create(const NamedValue &a, const NamedValue &b, const NamedValue &c) { return make({a, b, c}); }
};
inline SEXP wrap(const List &l) { return l.s; }

} /* namespace Rcpp */

#endif
