/*
 * ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Never linked into the product.
 *
 * Float64 C/OpenMP restatement of the hot path the reference (atoms-ufrj/atomsmm) delegates to
 * OpenMM: closed-form pair potentials over a cell-list-built Verlet list, bonded terms and the
 * RESPA step program with an optional Nose-Hoover (Suzuki-Yoshida) thermostat.  It is (a) the
 * large-system checker for the CUDA engine (sizes the numpy/sympy oracle cannot reach) and (b)
 * the CPU baseline timed by bench.py ("port": a restatement of the reference algorithm on the
 * host cores, not OpenMM itself, which is not available offline).
 *
 * Validated against oracle/refmath.py (generic evaluation of the energy strings, itself pinned
 * to the reference's goldens) in tests/test_oracle_c.py.
 *
 * Reference lines restated:
 *   near none/shift/force-switch      src/atomsmm/forces.py:541-563
 *   damped-smoothed                   src/atomsmm/forces.py:448-455
 *   LJ + reaction-field/erfc Coulomb  openmm.NonbondedForce as configured at forces.py:152-190
 *   exceptions                        src/atomsmm/forces.py:405-407
 *   RESPA recursion                   src/atomsmm/propagators.py:940-973
 *   Nose-Hoover                       src/atomsmm/propagators.py:1259-1273
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

enum { F_NEAR = 1, F_DAMPED = 2, F_LJC = 3, F_LJ_VIRIAL = 4 };
enum { B_BOND = 1, B_ANGLE = 2, B_TORSION = 3, B_LJC = 4 };

typedef struct {
    int family, group;
    double cutoff, p[16];
} orc_pair;

typedef struct {
    int family, group, nterms, arity, stride;
    int* atoms;
    double* params;
    double g[4];
} orc_bonded;

typedef struct {
    int n;
    double box[3];
    double *mass, *q, *sigma, *eps;
    int *excl_ptr, *excl_idx;
    int npair, nbonded;
    orc_pair pair[8];
    orc_bonded bonded[16];
    /* Verlet list (half list: j > i in storage order), one per distinct cutoff handled lazily */
    double list_cut, skin;
    int *nl_ptr, *nl_idx;
    double* xref;
    int threads;
    long rebuilds, pair_evals;
} orc_sys;

static double sw(double u) { return 1 - u*u*u*(10 - 15*u + 6*u*u); }
static double dsw(double u) { double o = 1 - u; return -30*u*u*o*o; }

/* E and rF = -r dE/dr of one pair */
static void pair_eval(const orc_pair* f, double r2, double qq, double sig, double eps, double* e, double* rF, int want_e) {
    const double* p = f->p;
    double r = sqrt(r2), s2 = sig*sig/r2, s6 = s2*s2*s2;
    double elj = 4*eps*s6*(s6 - 1), rflj = 24*eps*s6*(2*s6 - 1);
    *e = 0; *rF = 0;
    if (f->family == F_NEAR) {
        int variant = (int)p[0];
        double rs = p[1], rc = p[2], kc = p[5] != 0 ? p[3] : 0.0, sign = p[4];
        if (r >= rc) return;
        double u = r > rs ? (r - rs)/(rc - rs) : 0.0, S = sw(u), rdS = dsw(u)*r/(rc - rs);
        double ec = kc*qq/r, V = elj + ec, rf = rflj + ec;
        if (variant == 0) { *e = S*V; *rF = S*rf - rdS*V; }
        else if (variant == 1) {
            double c6 = pow(sig/rc, 6);
            V -= 4*eps*c6*(c6 - 1) + kc*qq/rc;
            *e = S*V; *rF = S*rf - rdS*V;
        } else if (!want_e) {
            *rF = S*rf;
        } else {
            double b = rs/(rc - rs), f12 = 1, f6 = 1, f1 = 1;
            if (u > 0) {
                double R = u/b + 1, u2 = u*u, u3 = u2*u, u4 = u2*u2, u5 = u4*u, b2 = b*b, b3 = b2*b;
                f12 += (6*b2 - 21*b + 28)*(b3*(pow(R, 12) - 1) - 12*b2*u - 66*b*u2 - 220*u3)/462 + 45*(7 - 2*b)*u4/14 - 72*u5/7;
                f6 += (6*b2 - 3*b + 1)*(b3*(pow(R, 6) - 1) - 6*b2*u - 15*b*u2 - 20*u3) + 45*(1 - 2*b)*u4 - 36*u5;
                f1 += 5*(b + 1)*(b + 1)*(6*b3*R*log(R) - 6*b2*u - 3*b*u2 + u3) + u4*(3*u - 5*b - 10)/2;
            }
            double f12c = pow(1 + b, 3)*(pow(b, 6) + 3*pow(b, 5) + (30.0/7)*pow(b, 4) + (25.0/7)*pow(b, 3) + (25.0/14)*b*b + 0.5*b + 2.0/33)/pow(b, 9);
            double f6c = pow(1 + b, 3)/pow(b, 3);
            double f1c = (30*(1 + b))*(b*b*(1 + b)*(1 + b)*log(1/b + 1) - b*b*b - 1.5*b*b - b/3 + 1.0/12);
            double c6 = pow(sig/rc, 6);
            *e = 4*eps*(f12*s6*s6 - f6*s6) + kc*qq*f1/r - (4*eps*(f12c*c6*c6 - f6c*c6) + kc*qq*f1c/rc);
            *rF = S*rf;
        }
        *e *= sign; *rF *= sign;
    } else if (f->family == F_DAMPED) {
        double alpha = p[0], rs = p[1], rc = p[2], kc = p[4];
        int d = (int)p[3];
        if (r >= rc) return;
        double er = erfc(alpha*r)/r;
        double ec = kc*qq*er, rfc = kc*qq*(er + 2*alpha/sqrt(M_PI)*exp(-alpha*alpha*r2));
        double u, rdu;
        if (d == 1) { u = r > rs ? (r - rs)/(rc - rs) : 0; rdu = r/(rc - rs); }
        else { double w = pow(rc, d) - pow(rs, d); u = r > rs ? (pow(r, d) - pow(rs, d))/w : 0; rdu = d*pow(r, d)/w; }
        double S = sw(u), rdS = dsw(u)*rdu, V = elj + ec;
        *e = S*V; *rF = S*(rflj + rfc) - rdS*V;
    } else if (f->family == F_LJC) {
        double kc = p[1] != 0 ? p[0] : 0.0, krf = p[2], crf = p[3], alpha = p[4], rs = p[6], rc = p[7];
        int kind = (int)p[1], use_switch = p[5] != 0;
        double S = 1, rdS = 0;
        if (use_switch && r > rs) { double u = (r - rs)/(rc - rs); S = sw(u); rdS = dsw(u)*r/(rc - rs); }
        double ec, rfc;
        if (kind == 2) { ec = kc*qq*(1/r + krf*r2 - crf); rfc = kc*qq*(1/r - 2*krf*r2); }
        else if (kind == 3) { double er = erfc(alpha*r)/r; ec = kc*qq*er; rfc = kc*qq*(er + 2*alpha/sqrt(M_PI)*exp(-alpha*alpha*r2)); }
        else { ec = kc*qq/r; rfc = ec; }
        *e = S*elj + ec; *rF = S*rflj - rdS*elj + rfc;
    } else if (f->family == F_LJ_VIRIAL) {
        double rs = p[1], rc = p[2], S = 1, rdS = 0;
        if (p[0] != 0 && r > rs) { double u = (r - rs)/(rc - rs); S = sw(u); rdS = dsw(u)*r/(rc - rs); }
        double V = 24*eps*s6*(2*s6 - 1), rf = 144*eps*s6*(4*s6 - 1);
        *e = S*V; *rF = S*rf - rdS*V;
    }
}

static double pair_range(const orc_pair* f) {
    if (f->family == F_NEAR || f->family == F_DAMPED) return f->p[2] < f->cutoff ? f->p[2] : f->cutoff;
    return f->cutoff;
}

/* ---- API ------------------------------------------------------------------------------------ */
int orc_create(int n, const double* mass, const double* box, orc_sys** out) {
    orc_sys* s = (orc_sys*)calloc(1, sizeof(orc_sys));
    s->n = n;
    memcpy(s->box, box, 3*sizeof(double));
    s->mass = (double*)malloc(n*sizeof(double));
    memcpy(s->mass, mass, n*sizeof(double));
    s->q = (double*)calloc(n, sizeof(double));
    s->sigma = (double*)calloc(n, sizeof(double));
    s->eps = (double*)calloc(n, sizeof(double));
    s->excl_ptr = (int*)calloc(n + 1, sizeof(int));
    s->excl_idx = (int*)calloc(1, sizeof(int));
    s->skin = 0.1;
    s->xref = (double*)malloc(3*n*sizeof(double));
#ifdef _OPENMP
    s->threads = omp_get_max_threads();
#else
    s->threads = 1;
#endif
    *out = s;
    return 0;
}

void orc_destroy(orc_sys* s) {
    if (!s) return;
    free(s->mass); free(s->q); free(s->sigma); free(s->eps); free(s->excl_ptr); free(s->excl_idx);
    free(s->nl_ptr); free(s->nl_idx); free(s->xref);
    for (int k = 0; k < s->nbonded; k++) { free(s->bonded[k].atoms); free(s->bonded[k].params); }
    free(s);
}

int orc_set_threads(orc_sys* s, int threads) {
    s->threads = threads > 0 ? threads : 1;
    return s->threads;
}

int orc_set_params(orc_sys* s, const double* q, const double* sigma, const double* eps) {
    memcpy(s->q, q, s->n*sizeof(double));
    memcpy(s->sigma, sigma, s->n*sizeof(double));
    memcpy(s->eps, eps, s->n*sizeof(double));
    return 0;
}

int orc_set_exclusions(orc_sys* s, int npairs, const int* pairs) {
    int n = s->n;
    memset(s->excl_ptr, 0, (n + 1)*sizeof(int));
    for (int k = 0; k < npairs; k++) { s->excl_ptr[pairs[2*k] + 1]++; s->excl_ptr[pairs[2*k+1] + 1]++; }
    for (int i = 0; i < n; i++) s->excl_ptr[i+1] += s->excl_ptr[i];
    free(s->excl_idx);
    s->excl_idx = (int*)malloc((2*npairs + 1)*sizeof(int));
    int* cur = (int*)malloc(n*sizeof(int));
    memcpy(cur, s->excl_ptr, n*sizeof(int));
    for (int k = 0; k < npairs; k++) {
        int i = pairs[2*k], j = pairs[2*k+1];
        s->excl_idx[cur[i]++] = j; s->excl_idx[cur[j]++] = i;
    }
    free(cur);
    free(s->nl_ptr); s->nl_ptr = NULL;
    return 0;
}

int orc_add_pair(orc_sys* s, int family, int group, double cutoff, const double* params, int nparams) {
    if (s->npair >= 8 || nparams > 16) return -1;
    orc_pair* f = &s->pair[s->npair++];
    f->family = family; f->group = group; f->cutoff = cutoff;
    memset(f->p, 0, sizeof(f->p));
    memcpy(f->p, params, nparams*sizeof(double));
    free(s->nl_ptr); s->nl_ptr = NULL;
    return 0;
}

int orc_add_bonded(orc_sys* s, int family, int group, int nterms, const int* atoms, const double* params,
                   int stride, const double* g, int ng) {
    if (s->nbonded >= 16) return -1;
    orc_bonded* b = &s->bonded[s->nbonded++];
    b->family = family; b->group = group; b->nterms = nterms; b->stride = stride;
    b->arity = family == B_ANGLE ? 3 : family == B_TORSION ? 4 : 2;
    b->atoms = (int*)malloc(sizeof(int)*b->arity*(nterms + 1));
    memcpy(b->atoms, atoms, sizeof(int)*b->arity*nterms);
    b->params = (double*)malloc(sizeof(double)*stride*(nterms + 1));
    memcpy(b->params, params, sizeof(double)*stride*nterms);
    for (int k = 0; k < 4; k++) b->g[k] = k < ng ? g[k] : 0.0;
    return 0;
}

static int excluded(const orc_sys* s, int i, int j) {
    for (int k = s->excl_ptr[i]; k < s->excl_ptr[i+1]; k++)
        if (s->excl_idx[k] == j) return 1;
    return 0;
}

/* cell-list build of a half Verlet list with radius rlist */
static void build_list(orc_sys* s, const double* x, double rlist) {
    int n = s->n, nc[3];
    double cs[3];
    for (int d = 0; d < 3; d++) { nc[d] = (int)floor(s->box[d]/rlist); if (nc[d] < 1) nc[d] = 1; cs[d] = s->box[d]/nc[d]; }
    int ncells = nc[0]*nc[1]*nc[2];
    int* head = (int*)malloc(ncells*sizeof(int));
    int* next = (int*)malloc(n*sizeof(int));
    int* cell = (int*)malloc(n*sizeof(int));
    for (int c = 0; c < ncells; c++) head[c] = -1;
    for (int i = n - 1; i >= 0; i--) {
        int c[3];
        for (int d = 0; d < 3; d++) {
            double w = x[3*i+d] - s->box[d]*floor(x[3*i+d]/s->box[d]);
            c[d] = (int)(w/cs[d]); if (c[d] >= nc[d]) c[d] = nc[d] - 1; if (c[d] < 0) c[d] = 0;
        }
        cell[i] = (c[2]*nc[1] + c[1])*nc[0] + c[0];
        next[i] = head[cell[i]]; head[cell[i]] = i;
    }
    free(s->nl_ptr); free(s->nl_idx);
    s->nl_ptr = (int*)malloc((n + 1)*sizeof(int));
    int* count = (int*)calloc(n, sizeof(int));
    double r2max = rlist*rlist;
    int** rows = (int**)malloc(n*sizeof(int*));
#pragma omp parallel for schedule(dynamic, 64) num_threads(s->threads)
    for (int i = 0; i < n; i++) {
        int cap = 256, m = 0;
        int* row = (int*)malloc(cap*sizeof(int));
        int ci = cell[i], cx = ci % nc[0], cy = (ci/nc[0]) % nc[1], cz = ci/(nc[0]*nc[1]);
        int seen[27], nseen = 0;
        for (int dz = -1; dz <= 1; dz++) for (int dy = -1; dy <= 1; dy++) for (int dx = -1; dx <= 1; dx++) {
            int c = (((cz + dz + nc[2]) % nc[2])*nc[1] + (cy + dy + nc[1]) % nc[1])*nc[0] + (cx + dx + nc[0]) % nc[0];
            int dup = 0;
            for (int k = 0; k < nseen; k++) if (seen[k] == c) dup = 1;
            if (dup) continue;
            seen[nseen++] = c;
            for (int j = head[c]; j >= 0; j = next[j]) {
                if (j <= i) continue;
                double r2 = 0;
                for (int d = 0; d < 3; d++) { double t = x[3*j+d] - x[3*i+d]; t -= s->box[d]*rint(t/s->box[d]); r2 += t*t; }
                if (r2 < r2max && !excluded(s, i, j)) {
                    if (m == cap) { cap *= 2; row = (int*)realloc(row, cap*sizeof(int)); }
                    row[m++] = j;
                }
            }
        }
        rows[i] = row; count[i] = m;
    }
    s->nl_ptr[0] = 0;
    for (int i = 0; i < n; i++) s->nl_ptr[i+1] = s->nl_ptr[i] + count[i];
    s->nl_idx = (int*)malloc((s->nl_ptr[n] + 1)*sizeof(int));
    for (int i = 0; i < n; i++) { memcpy(s->nl_idx + s->nl_ptr[i], rows[i], count[i]*sizeof(int)); free(rows[i]); }
    free(rows); free(count); free(head); free(next); free(cell);
    memcpy(s->xref, x, 3*n*sizeof(double));
    s->list_cut = rlist;
    s->rebuilds++;
}

static void ensure_list(orc_sys* s, const double* x) {
    double rmax = 0;
    for (int k = 0; k < s->npair; k++) { double r = pair_range(&s->pair[k]); if (r > rmax) rmax = r; }
    if (rmax == 0) return;
    double rlist = rmax + s->skin;
    int need = s->nl_ptr == NULL || fabs(s->list_cut - rlist) > 1e-12;
    if (!need) {
        double lim = 0.25*s->skin*s->skin;
        for (int i = 0; i < s->n && !need; i++) {
            double d2 = 0;
            for (int d = 0; d < 3; d++) { double t = x[3*i+d] - s->xref[3*i+d]; d2 += t*t; }
            if (d2 > lim) need = 1;
        }
    }
    if (need) build_list(s, x, rlist);
}

static void bonded_eval(const orc_sys* s, const orc_bonded* b, const double* x, double* f, double* e_out, double* w_out) {
    double e_tot = 0, w_tot = 0;
    for (int t = 0; t < b->nterms; t++) {
        const int* a = b->atoms + t*b->arity;
        const double* p = b->params + (size_t)t*b->stride;
        if (b->arity == 2) {
            double d[3], r2 = 0;
            for (int k = 0; k < 3; k++) { d[k] = x[3*a[1]+k] - x[3*a[0]+k]; d[k] -= s->box[k]*rint(d[k]/s->box[k]); r2 += d[k]*d[k]; }
            double r = sqrt(r2), e, dedr;
            if (b->family == B_BOND) { e = 0.5*p[1]*(r - p[0])*(r - p[0]); dedr = p[1]*(r - p[0]); }
            else {
                double s2 = p[1]*p[1]/r2, s6 = s2*s2*s2;
                e = 4*p[2]*s6*(s6 - 1) + b->g[0]*p[0]/r;
                dedr = -(24*p[2]*s6*(2*s6 - 1) + b->g[0]*p[0]/r)/r;
                if (b->g[1] > 0) {
                    double al = b->g[1], kq = b->g[0]*p[3], er = erf(al*r);
                    e -= kq*er/r;
                    dedr -= kq*(2*al/sqrt(M_PI)*exp(-al*al*r2)/r - er/r2);
                }
            }
            e_tot += e; w_tot += -dedr*r;
            if (f) for (int k = 0; k < 3; k++) { double g = dedr/r*d[k]; f[3*a[0]+k] += g; f[3*a[1]+k] -= g; }
        } else if (b->arity == 3) {
            double u[3], v[3], ru2 = 0, rv2 = 0, uv = 0;
            for (int k = 0; k < 3; k++) { u[k] = x[3*a[0]+k] - x[3*a[1]+k]; v[k] = x[3*a[2]+k] - x[3*a[1]+k]; ru2 += u[k]*u[k]; rv2 += v[k]*v[k]; uv += u[k]*v[k]; }
            double ru = sqrt(ru2), rv = sqrt(rv2), c = uv/(ru*rv);
            if (c > 1) c = 1; if (c < -1) c = -1;
            double th = acos(c), dedt = p[1]*(th - p[0]), sn = sqrt(fmax(1 - c*c, 1e-30));
            e_tot += 0.5*p[1]*(th - p[0])*(th - p[0]);
            if (f) for (int k = 0; k < 3; k++) {
                double fi = dedt*(v[k]/(ru*rv) - c*u[k]/ru2)/sn, fk = dedt*(u[k]/(ru*rv) - c*v[k]/rv2)/sn;
                f[3*a[0]+k] += fi; f[3*a[2]+k] += fk; f[3*a[1]+k] -= fi + fk;
            }
        } else {
            double F[3], G[3], H[3], A[3], B[3], C[3];
            for (int k = 0; k < 3; k++) { F[k] = x[3*a[0]+k] - x[3*a[1]+k]; G[k] = x[3*a[1]+k] - x[3*a[2]+k]; H[k] = x[3*a[3]+k] - x[3*a[2]+k]; }
            A[0] = F[1]*G[2] - F[2]*G[1]; A[1] = F[2]*G[0] - F[0]*G[2]; A[2] = F[0]*G[1] - F[1]*G[0];
            B[0] = H[1]*G[2] - H[2]*G[1]; B[1] = H[2]*G[0] - H[0]*G[2]; B[2] = H[0]*G[1] - H[1]*G[0];
            C[0] = B[1]*A[2] - B[2]*A[1]; C[1] = B[2]*A[0] - B[0]*A[2]; C[2] = B[0]*A[1] - B[1]*A[0];
            double A2 = A[0]*A[0] + A[1]*A[1] + A[2]*A[2], B2 = B[0]*B[0] + B[1]*B[1] + B[2]*B[2];
            double G2 = G[0]*G[0] + G[1]*G[1] + G[2]*G[2], gn = sqrt(G2), nrm = sqrt(A2*B2);
            double cp = (A[0]*B[0] + A[1]*B[1] + A[2]*B[2])/nrm, sp = (C[0]*G[0] + C[1]*G[1] + C[2]*G[2])/(nrm*gn);
            double phi = atan2(sp, cp);
            e_tot += p[2]*(1 + cos(p[0]*phi - p[1]));
            if (f) {
                double dedphi = -p[2]*p[0]*sin(p[0]*phi - p[1]);
                double fg = F[0]*G[0] + F[1]*G[1] + F[2]*G[2], hg = H[0]*G[0] + H[1]*G[1] + H[2]*G[2];
                for (int k = 0; k < 3; k++) {
                    double d1 = -gn/A2*A[k], d4 = gn/B2*B[k];
                    double d2 = gn/A2*A[k] + fg/(A2*gn)*A[k] - hg/(B2*gn)*B[k];
                    double d3 = -gn/B2*B[k] - fg/(A2*gn)*A[k] + hg/(B2*gn)*B[k];
                    f[3*a[0]+k] -= dedphi*d1; f[3*a[1]+k] -= dedphi*d2; f[3*a[2]+k] -= dedphi*d3; f[3*a[3]+k] -= dedphi*d4;
                }
            }
        }
    }
    *e_out += e_tot; *w_out += w_tot;
}

/* forces (kJ/mol/nm) of the groups in mask are ADDED into f (caller zeroes); energy/virial summed */
int orc_eval(orc_sys* s, const double* x, unsigned mask, double* f, double* energy, double* virial) {
    int n = s->n;
    const int want_e = energy != NULL;
    double e_tot = 0, w_tot = 0;
    int any = 0;
    for (int k = 0; k < s->npair; k++) if (mask & (1u << s->pair[k].group)) any = 1;
    if (any) {
        ensure_list(s, x);
        int T = s->threads;
        double* priv = f ? (double*)calloc((size_t)T*3*n, sizeof(double)) : NULL;
#pragma omp parallel num_threads(T) reduction(+:e_tot, w_tot)
        {
#ifdef _OPENMP
            int tid = omp_get_thread_num();
#else
            int tid = 0;
#endif
            double* fp = priv ? priv + (size_t)tid*3*n : NULL;
#pragma omp for schedule(dynamic, 128)
            for (int i = 0; i < n; i++) {
                double fi[3] = {0, 0, 0};
                for (int m = s->nl_ptr[i]; m < s->nl_ptr[i+1]; m++) {
                    int j = s->nl_idx[m];
                    double d[3], r2 = 0;
                    for (int k = 0; k < 3; k++) { d[k] = x[3*i+k] - x[3*j+k]; d[k] -= s->box[k]*rint(d[k]/s->box[k]); r2 += d[k]*d[k]; }
                    double qq = s->q[i]*s->q[j], sig = 0.5*(s->sigma[i] + s->sigma[j]), eps = sqrt(s->eps[i]*s->eps[j]);
                    double rF_sum = 0;
                    for (int k = 0; k < s->npair; k++) {
                        const orc_pair* pf = &s->pair[k];
                        if (!(mask & (1u << pf->group))) continue;
                        double rc = pair_range(pf);
                        if (r2 >= rc*rc) continue;
                        double e, rF;
                        pair_eval(pf, r2, qq, sig, eps, &e, &rF, want_e);
                        e_tot += e; w_tot += rF; rF_sum += rF;
                    }
                    if (fp && rF_sum != 0) {
                        double g = rF_sum/r2;
                        for (int k = 0; k < 3; k++) { fi[k] += g*d[k]; fp[3*j+k] -= g*d[k]; }
                    }
                }
                if (fp) for (int k = 0; k < 3; k++) fp[3*i+k] += fi[k];
            }
        }
        s->pair_evals++;
        if (f) {
#pragma omp parallel for num_threads(T)
            for (int d = 0; d < 3*n; d++) { double t = 0; for (int k = 0; k < T; k++) t += priv[(size_t)k*3*n + d]; f[d] += t; }
            free(priv);
        }
    }
    for (int k = 0; k < s->nbonded; k++)
        if (mask & (1u << s->bonded[k].group)) bonded_eval(s, &s->bonded[k], x, f, &e_tot, &w_tot);
    if (energy) *energy = e_tot;
    if (virial) *virial = w_tot;
    return 0;
}

/* Exact interacting pair set of pair force `which`: i < j, r^2 < rc^2 (float64, sum order x,y,z), not
 * excluded -> number of pairs and an order-independent checksum (sum of a 64-bit mix of (i<<32|j)), the same
 * definition as the engine's k_pair_set: "neighbour lists bit-exact" is checked as equality of both. */
static unsigned long long mix64(unsigned long long z) {
    z += 0x9e3779b97f4a7c15ull;
    z = (z ^ (z >> 30))*0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27))*0x94d049bb133111ebull;
    return z ^ (z >> 31);
}

int orc_pair_set(orc_sys* s, const double* x, int which, long long* count, unsigned long long* checksum) {
    if (which < 0 || which >= s->npair) return -1;
    ensure_list(s, x);
    const double rc = pair_range(&s->pair[which]), rc2 = rc*rc;
    long long c = 0;
    unsigned long long h = 0;
#pragma omp parallel for schedule(dynamic, 128) num_threads(s->threads) reduction(+:c, h)
    for (int i = 0; i < s->n; i++) {
        for (int m = s->nl_ptr[i]; m < s->nl_ptr[i+1]; m++) {
            int j = s->nl_idx[m];
            double r2 = 0;
            for (int k = 0; k < 3; k++) { double d = x[3*j+k] - x[3*i+k]; d -= s->box[k]*rint(d/s->box[k]); r2 += d*d; }
            if (r2 < rc2) { c++; h += mix64(((unsigned long long)i << 32) | (unsigned)j); }
        }
    }
    *count = c;
    *checksum = h;
    return 0;
}

static void kick(int n, double* v, const double* f, const double* g, const double* mass, double c) {
    for (int i = 0; i < n; i++) if (mass[i] > 0) for (int k = 0; k < 3; k++) v[3*i+k] += c*(f[3*i+k] - (g ? g[3*i+k] : 0.0))/mass[i];
}

static double mvv(int n, const double* v, const double* mass) {
    double s = 0;
    for (int i = 0; i < n; i++) s += mass[i]*(v[3*i]*v[3*i] + v[3*i+1]*v[3*i+1] + v[3*i+2]*v[3*i+2]);
    return s;
}

/* NoseHooverPropagator.addSteps with nloops (incl. the n > 2 sub-loop guard) for a fraction of dt */
static void nose_hoover(int n, double* v, const double* mass, double h_total, int nloops, double LkT, double Q, double* p_eta) {
    double h = h_total/nloops, m2 = mvv(n, v, mass);
    *p_eta += 0.5*h*(m2 - LkT);
    double vs = exp(-h*(*p_eta)/Q);
    if (nloops > 2) for (int k = 1; k < nloops; k++) { *p_eta += h*(vs*vs*m2 - LkT); vs *= exp(-h*(*p_eta)/Q); }
    *p_eta += 0.5*h*(vs*vs*m2 - LkT);
    for (int d = 0; d < 3*n; d++) v[d] *= vs;
}

/* RespaPropagator([n0, n1, 1]) with groups 0/1/2 and kick (f2 - f1) at the outer level, optionally
 * wrapped as TrotterSuzuki(Respa, SuzukiYoshida(NoseHoover(nloops), 3)).  nh: 0 none, 1 on. */
int orc_respa(orc_sys* s, double* x, double* v, int nsteps, double dt, int n0, int n1, int nh, int nloops,
              double LkT, double Q, double* p_eta) {
    int n = s->n;
    double *f0 = (double*)malloc(3*n*sizeof(double)), *f1 = (double*)malloc(3*n*sizeof(double)), *f2 = (double*)malloc(3*n*sizeof(double));
    const double w[3] = {1.3512071919596578, 1 - 2*1.3512071919596578, 1.3512071919596578};
    memset(f1, 0, 3*n*sizeof(double)); memset(f2, 0, 3*n*sizeof(double));
    orc_eval(s, x, 1u << 1, f1, NULL, NULL);
    orc_eval(s, x, 1u << 2, f2, NULL, NULL);
    for (int step = 0; step < nsteps; step++) {
        if (nh) for (int k = 0; k < 3; k++) nose_hoover(n, v, s->mass, 0.5*w[k]*dt, nloops, LkT, Q, p_eta);
        kick(n, v, f2, f1, s->mass, 0.5*dt);
        for (int a = 0; a < n1; a++) {
            double h1 = dt/n1;
            kick(n, v, f1, NULL, s->mass, 0.5*h1);
            for (int b = 0; b < n0; b++) {
                double h0 = h1/n0;
                memset(f0, 0, 3*n*sizeof(double));
                orc_eval(s, x, 1u, f0, NULL, NULL);
                kick(n, v, f0, NULL, s->mass, 0.5*h0);
                for (int i = 0; i < n; i++) if (s->mass[i] > 0) for (int k = 0; k < 3; k++) x[3*i+k] += h0*v[3*i+k];
                memset(f0, 0, 3*n*sizeof(double));
                orc_eval(s, x, 1u, f0, NULL, NULL);
                kick(n, v, f0, NULL, s->mass, 0.5*h0);
            }
            memset(f1, 0, 3*n*sizeof(double));
            orc_eval(s, x, 1u << 1, f1, NULL, NULL);
            kick(n, v, f1, NULL, s->mass, 0.5*h1);
        }
        memset(f2, 0, 3*n*sizeof(double));
        orc_eval(s, x, 1u << 2, f2, NULL, NULL);
        kick(n, v, f2, f1, s->mass, 0.5*dt);
        if (nh) for (int k = 0; k < 3; k++) nose_hoover(n, v, s->mass, 0.5*w[k]*dt, nloops, LkT, Q, p_eta);
    }
    free(f0); free(f1); free(f2);
    return 0;
}

long orc_counter(const orc_sys* s, int which) { return which == 0 ? s->rebuilds : s->pair_evals; }
