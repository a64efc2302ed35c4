// pw_lsoda.cuh -- LSODA (Petzold & Hindmarsh, ODEPACK) restated for small systems,
// used where the reference calls `scipy.integrate.odeint` (helium.py:694, :736).
//
// scipy's odeint drives ODEPACK's LSODA (scipy/integrate/_odepack: lsoda, stoda,
// cfode, prja, solsy, intdy, ewset, vmnorm, fnorm) with itask = 1, itol = 1,
// rtol = atol = 1.49012e-8, jt = 2 (internally generated full Jacobian),
// mxstep = 500, mxordn = 12, mxords = 5, h0 = hmax = hmin = 0.  No LSODA source
// exists in this image (only scipy's compiled extension), so this is a
// restatement of the published algorithm: Nordsieck history array, variable-order
// Adams (1..12) / BDF (1..5) with automatic stiffness switching, chord iteration
// with a finite-difference Jacobian, step/order selection from the three error
// estimates, output by interpolation (intdy).  tests/test_hostsim_solvers.py holds
// the host build to scipy's own `odeint(full_output=True)` step/feval/jacobian
// counters and to its solution on the anchor models.
#pragma once
#include "pw_common.cuh"

#if defined(PW_LSODA_HIST) && !defined(__CUDA_ARCH__)
extern "C" long pw_lsoda_hist[2][13];  // test instrumentation: successful steps per (method, order)
long pw_lsoda_hist[2][13];
#endif

namespace pw {

constexpr int kLsodaMaxOrdN = 12;  // Adams
constexpr int kLsodaMaxOrdS = 5;   // BDF

// Method coefficients (ODEPACK cfode): elco[meth][nq][i], tesco[meth][nq][k], 1-based.
struct LsodaTables {
  double elco[2][13][14];
  double tesco[2][13][4];
  double rtesco2[2][13];  // 1 / tesco[meth][nq][2] (label 700 scales acor by it)
  double rtesco1[2][13], rtesco3[2][13];  // 1 / tesco[meth][nq][1], [3] (PW_RDIV)
  double conit[13];     // 0.5 / (nq + 2): stoda label 150
  double rtcon[2][13];  // 1 / (tesco[meth][nq][2] * conit[nq]) (PW_RDIV in the corrector's convergence test)
  double cm1[13], cm2[6];
  double sm1[13];  // Adams stability-region bounds used by the switching logic
};

PW_HD void lsoda_tables_init(LsodaTables* t) {
  for (int m = 0; m < 2; ++m)
    for (int a = 0; a < 13; ++a) {
      for (int b = 0; b < 14; ++b) t->elco[m][a][b] = 0.0;
      for (int b = 0; b < 4; ++b) t->tesco[m][a][b] = 0.0;
    }
  double pc[13];
  {  // meth = 1: Adams-Moulton, orders 1..12
    double (*elco)[14] = t->elco[0];
    double (*tesco)[4] = t->tesco[0];
    elco[1][1] = 1.; elco[1][2] = 1.;
    tesco[1][1] = 0.; tesco[1][2] = 2.; tesco[2][1] = 1.; tesco[12][3] = 0.;
    pc[1] = 1.;
    double rqfac = 1.;
    for (int nq = 2; nq <= 12; ++nq) {
      const double rq1fac = rqfac;
      rqfac = rqfac / (double)nq;
      const int nqm1 = nq - 1, nqp1 = nq + 1;
      const double fnqm1 = (double)nqm1;
      pc[nq] = 0.;
      for (int i = nq; i >= 2; --i) pc[i] = pc[i - 1] + fnqm1 * pc[i];
      pc[1] = fnqm1 * pc[1];
      double pint = pc[1], xpin = pc[1] / 2., tsign = 1.;
      for (int i = 2; i <= nq; ++i) {
        tsign = -tsign;
        pint += tsign * pc[i] / (double)i;
        xpin += tsign * pc[i] / (double)(i + 1);
      }
      elco[nq][1] = pint * rq1fac;
      elco[nq][2] = 1.;
      for (int i = 2; i <= nq; ++i) elco[nq][i + 1] = rq1fac * pc[i] / (double)i;
      const double agamq = rqfac * xpin;
      const double ragq = 1. / agamq;
      tesco[nq][2] = ragq;
      if (nq < 12) tesco[nqp1][1] = ragq * rqfac / (double)nqp1;
      tesco[nqm1][3] = ragq;
    }
  }
  {  // meth = 2: BDF, orders 1..5
    double (*elco)[14] = t->elco[1];
    double (*tesco)[4] = t->tesco[1];
    pc[1] = 1.;
    double rq1fac = 1.;
    for (int nq = 1; nq <= 5; ++nq) {
      const double fnq = (double)nq;
      const int nqp1 = nq + 1;
      pc[nqp1] = 0.;
      for (int i = nq + 1; i >= 2; --i) pc[i] = pc[i - 1] + fnq * pc[i];
      pc[1] *= fnq;
      for (int i = 1; i <= nqp1; ++i) elco[nq][i] = pc[i] / pc[2];
      elco[nq][2] = 1.;
      tesco[nq][1] = rq1fac;
      tesco[nq][2] = ((double)nqp1) / elco[nq][1];
      tesco[nq][3] = ((double)(nq + 2)) / elco[nq][1];
      rq1fac /= fnq;
    }
  }
  for (int m = 0; m < 2; ++m)
    for (int a = 0; a < 13; ++a) {
      t->rtesco2[m][a] = (t->tesco[m][a][2] != 0.) ? 1. / t->tesco[m][a][2] : 0.;
      t->rtesco1[m][a] = (t->tesco[m][a][1] != 0.) ? 1. / t->tesco[m][a][1] : 0.;
      t->rtesco3[m][a] = (t->tesco[m][a][3] != 0.) ? 1. / t->tesco[m][a][3] : 0.;
      t->conit[a] = 0.5 / (double)(a + 2);
      t->rtcon[m][a] = (t->tesco[m][a][2] != 0.) ? 1. / (t->tesco[m][a][2] * t->conit[a]) : 0.;
    }
  for (int i = 1; i <= 5; ++i) t->cm2[i] = t->tesco[1][i][2] * t->elco[1][i][i + 1];
  for (int i = 1; i <= 12; ++i) t->cm1[i] = t->tesco[0][i][2] * t->elco[0][i][i + 1];
  const double s[13] = {0., 0.5, 0.575, 0.55, 0.45, 0.35, 0.25, 0.2, 0.15, 0.1, 0.075, 0.05, 0.025};
  for (int i = 0; i < 13; ++i) t->sm1[i] = s[i];
}

struct LsodaOpts {
  double rtol, atol;
  int mxstep;   // per output interval (odeint default 500)
  double hmax;  // <= 0: none
  double h0;    // <= 0: automatic
};

// Run-time rows of the Nordsieck array (orders above kF; see Lsoda below): all passes of the
// Pascal-triangle product that touch rows kF+2..nq+1.  hand[(kF+1-j)*N + i] receives row
// kF+2 as it stands at the start of pass j (j = kF+1..1), which is what pass j adds to row kF+1.
template <int N>
PW_HD_NOINLINE void lsoda_pascal_hi(double* yhs, int nq, int kF, double sign, double* hand) {
  for (int j = nq; j >= kF + 2; --j)
    for (int i1 = j; i1 <= nq; ++i1)
      for (int i = 0; i < N; ++i) yhs[i1 * N + i] += sign * yhs[(i1 + 1) * N + i];
  for (int j = kF + 1; j >= 1; --j) {
    for (int i = 0; i < N; ++i) hand[(kF + 1 - j) * N + i] = yhs[(kF + 2) * N + i];
    for (int i1 = kF + 2; i1 <= nq; ++i1)
      for (int i = 0; i < N; ++i) yhs[i1 * N + i] += sign * yhs[(i1 + 1) * N + i];
  }
}

// istate values on return from call(): 2 ok, < 0 failure (ODEPACK numbering)
//  -1 excess work (mxstep), -2 excess accuracy, -3 illegal input, -4 repeated
//  error-test failures, -5 repeated convergence failures, -6 ewt <= 0
//
// Register-resident layout.  One thread integrates one model and the solve is a chain
// of ~10^4 dependent steps, so what matters is the latency of one step.  The Nordsieck
// rows 1..kF+1 (orders <= kF = 5: every BDF order, and the Adams orders a stiff problem
// ever reaches before the switch to BDF) are addressed only through loops that run to a
// compile-time bound under an `if (row <= nq)` guard and are fully unrolled: all indices
// into yh / wm / ipvt are constants, the arrays live in registers, and a guarded-off row
// costs one issue slot instead of a local-memory round trip.  The rows above (Adams
// orders 6..12, and ODEPACK's scratch row lmax = 13 of the Adams method) sit in `yhs`
// and are walked by ordinary run-time loops behind `if (nq > kF)`.  Rows addressed by a
// run-time index (yh[l], yh[lmax]) go through row_get / row_set.  The method coefficients are not copied (ODEPACK's el array): `elp`
// points at the row of the coefficient table that dcfode would have loaded, and the three
// test constants of the current order are cached when that row is (re)selected.
//
// RHS concept:  f.at(t)  makes t the current abscissa (may cache t-dependent terms);
//               f.eval(y, dydt)  evaluates at the current abscissa.
template <int N, class RHS>
struct Lsoda {
  static constexpr int kQ = kLsodaMaxOrdN;  // highest order; rows 1..kQ+1 of the history are used
  static constexpr int kF = kLsodaMaxOrdS;  // orders whose rows 1..kF+1 are register resident
  RHS& f;
  const LsodaTables& tab;
  LsodaOpts o;
  double yh[kF + 2][N];   // Nordsieck array, rows 1..kF+1
  // rows kF+2..kQ+1, run-time indexed: caller-provided storage of kHiDoubles doubles, kept
  // outside this object so that the object itself is only ever addressed by constants
  static constexpr int kHiDoubles = (kQ + 2) * N;
  double* yhs_;
  double y[N], ewt[N], savf[N], acor[N];
  double ewi[N];  // rtol*|y| + atol, the quantity ewt is the reciprocal of (PW_RDIV)
  double wm[N][N];
  int ipvt[N];
  const double* elp;     // elco[meth_tab][nq] in force (ODEPACK: el(1..l))
  double tq1, tq2, tq3;  // tesco(nq, 1..3) of the same table row
  double rtq2, rtq2u;    // 1 / tesco(nq, 2); the same for (nqu, table in force) as label 700 reads it
  double rtq1, rtq3, rtc;  // 1 / tesco(nq, 1), 1 / tesco(nq, 3), 1 / (tesco(nq, 2) * conit)   (PW_RDIV)
  double rdiag[N];       // 1 / diagonal of the LU factors (PW_RDIV in solsy)
  double tn, h, hu, hold, rc, el0, crate, conit, rmax, pdnorm, pdest, pdlast, ratio, hmxi, tsw;
  int nq, l, lmax, meth, mused, meth_tab, miter, jtyp, maxord, mxordn, mxords;
  int ialth, ipup, nslp, icount, irflag, jstart, kflag, jcur, ierpj, iersl;
  int nst, nfe, nje, nqu, nslast, nhnil;
  bool started;

  PW_HD Lsoda(RHS& f_, const LsodaTables& t_, const LsodaOpts& o_, double* hi_rows)
      : f(f_), tab(t_), o(o_), yhs_(hi_rows), started(false) {}

  PW_HD double vmnorm(const double* v) const {
    double vm = pw_pos_or_zero(fabs(v[0]) * ewt[0]);
#pragma unroll
    for (int i = 1; i < N; ++i) {
      const double a = fabs(v[i]) * ewt[i];
      vm = (a > vm) ? a : vm;  // == fmax(vm, a) for vm >= 0
    }
    return vm;
  }
  PW_HD bool ewset(const double* yc) {
    bool ok = true;
#pragma unroll
    for (int i = 0; i < N; ++i) {
      const double e = o.rtol * fabs(yc[i]) + o.atol;
      if (!(e > 0.0)) ok = false;
      ewi[i] = e;
      ewt[i] = PW_DIV(1.0, e);
    }
    return ok;
  }
  // yh[k] for a run-time row index
  PW_HD void row_get(int k, double* out) const {
    if (PW_UNLIKELY(k > kF + 1)) {
      for (int i = 0; i < N; ++i) out[i] = yhs_[(k) * N + i];
      return;
    }
    // every row is read unconditionally and the *value* is selected: a conditional read
    // per row would be merged by the compiler into one read at a run-time address, which
    // pins the whole object to local memory
#pragma unroll
    for (int i = 0; i < N; ++i) out[i] = yh[1][i];
#pragma unroll
    for (int j = 2; j <= kF + 1; ++j) {
#pragma unroll
      for (int i = 0; i < N; ++i) {
        const double v = yh[j][i];
        out[i] = (j == k) ? v : out[i];
      }
    }
  }
  PW_HD void row_set(int k, const double* in) {
    if (PW_UNLIKELY(k > kF + 1)) {
      for (int i = 0; i < N; ++i) yhs_[(k) * N + i] = in[i];
      return;
    }
#pragma unroll
    for (int j = 1; j <= kF + 1; ++j) {
#pragma unroll
      for (int i = 0; i < N; ++i) {
        const double v = yh[j][i];
        yh[j][i] = (j == k) ? in[i] : v;
      }
    }
  }
  // ODEPACK keeps ONE elco/tesco table in common storage and reloads it (cfode) only when
  // stoda is re-entered with jstart = -1 after a method switch; until then the old
  // method's coefficients stay in force (e.g. tesco(2,nqu) at label 700).
  PW_HD double sm1(int q) const { return tab.sm1[q]; }

  PW_HD void resetcoeff() {  // stoda label 150
    elp = tab.elco[meth_tab - 1][nq];
    const double* tq = tab.tesco[meth_tab - 1][nq];
    tq1 = tq[1]; tq2 = tq[2]; tq3 = tq[3];
    rtq2 = tab.rtesco2[meth_tab - 1][nq];
    rtq1 = tab.rtesco1[meth_tab - 1][nq];
    rtq3 = tab.rtesco3[meth_tab - 1][nq];
    const double el1 = elp[1];
    rc = PW_DIV(rc * el1, el0);
    el0 = el1;
    conit = tab.conit[nq];
    rtc = tab.rtcon[meth_tab - 1][nq];
  }
  PW_HD void scaleh(double* rh, double* pdh) {  // stoda labels 175-180
    *rh = pw_min(*rh, rmax);
    if (hmxi != 0.) *rh = PW_DIV(*rh, pw_max(1., fabs(h) * hmxi * (*rh)));  // (hmxi = 0: rh / 1)
    if (meth == 1) {
      irflag = 0;
      *pdh = pw_max(fabs(h) * pdlast, 0.000001);
      if ((*rh * *pdh * 1.00001) >= sm1(nq)) {
        *rh = PW_DIV(sm1(nq), *pdh);
        irflag = 1;
      }
    }
    double r = 1.;
#pragma unroll
    for (int j = 2; j <= kF + 1; ++j)
      if (j <= l) {
        r *= *rh;
#pragma unroll
        for (int i = 0; i < N; ++i) yh[j][i] *= r;
      }
    if (PW_UNLIKELY(nq > kF))
#pragma unroll 1
      for (int j = kF + 2; j <= l; ++j) {
        r *= *rh;
        for (int i = 0; i < N; ++i) yhs_[(j) * N + i] *= r;
      }
    h *= *rh;
    rc *= *rh;
    ialth = l;
  }
  // yh <- yh * Pascal triangle (sgn = +1, prediction) or its inverse (sgn = -1, retraction after
  // a failed step).  Each pass j = nq..1 adds row i1+1 to row i1 for i1 = j..nq in ascending
  // order.  The sign is a run-time +-1.0 and stoda() has a single call site, so that prediction
  // and retraction share ONE copy of the code (a +- b == fma(+-1, b, a) bit for bit): a second
  // copy would sit in the middle of the step loop's instruction-cache footprint.  For orders
  // above kF the run-time rows are advanced out of line first (lsoda_pascal_hi); what the
  // register rows need from them -- row kF+2 as it stands before each of the last kF+1
  // passes -- comes back in `hand`.
  PW_HD void pascal(double sgn) {
    // one copy for all register-resident orders: the row operations of order kF, each under
    // the guard `row <= nq` (a guarded-off operation costs its issue slot)
    if (nq <= kF) {
#pragma unroll
      for (int j = kF; j >= 1; --j)
#pragma unroll
        for (int i1 = j; i1 <= kF; ++i1) {
          const bool on = (i1 <= nq);
#pragma unroll
          for (int i = 0; i < N; ++i)
            if (on) yh[i1][i] = pw_fma(sgn, yh[i1 + 1][i], yh[i1][i]);
        }
      return;
    }
    double hand[(kF + 1) * N];
    lsoda_pascal_hi<N>(yhs_, nq, kF, sgn, hand);
#pragma unroll
    for (int j = kF + 1; j >= 1; --j) {
#pragma unroll
      for (int i1 = j; i1 <= kF; ++i1) {
#pragma unroll
        for (int i = 0; i < N; ++i) yh[i1][i] = pw_fma(sgn, yh[i1 + 1][i], yh[i1][i]);
      }
#pragma unroll
      for (int i = 0; i < N; ++i) yh[kF + 1][i] = pw_fma(sgn, hand[(kF + 1 - j) * N + i], yh[kF + 1][i]);
    }
  }
  // prja, miter = 2: finite-difference Jacobian, P = I - h*el0*J, LU (dgefa)
  PW_HD void prja() {
    ++nje;
    ierpj = 0;
    jcur = 1;
    const double hl0 = h * el0;
    double fac = vmnorm(savf);
    double r0 = 1000. * fabs(h) * kEps * ((double)N) * fac;
    if (r0 == 0.) r0 = 1.;
    const double srur = sqrt(kEps);
    // one column per pass of a *rolled* loop (a single copy of the right-hand side in the
    // code); the column index only ever selects values, never addresses
#pragma unroll 1
    for (int j = 0; j < N; ++j) {
      double yj = 0., ej = 0., eij = 0.;
#pragma unroll
      for (int i = 0; i < N; ++i) {
        yj = (i == j) ? y[i] : yj;
        ej = (i == j) ? ewt[i] : ej;
        eij = (i == j) ? ewi[i] : eij;
      }
      const double r = pw_max(srur * fabs(yj), PW_RDIV(r0, ej, eij));
      double yp[N], ftem[N];
#pragma unroll
      for (int i = 0; i < N; ++i) yp[i] = (i == j) ? yj + r : y[i];
      fac = PW_DIV(-hl0, r);
      f.eval(yp, ftem);
#pragma unroll
      for (int i = 0; i < N; ++i) {
        const double d = (ftem[i] - savf[i]) * fac;
#pragma unroll
        for (int jj = 0; jj < N; ++jj) wm[i][jj] = (jj == j) ? d : wm[i][jj];
      }
    }
    nfe += N;
    // pdnorm = fnorm(wm, ewt) / |hl0| ; fnorm = max_i w_i sum_j |a_ij| / w_j
    double an = 0.;
#pragma unroll
    for (int i = 0; i < N; ++i) {
      double sum = 0.;
#pragma unroll
      for (int j = 0; j < N; ++j) sum += PW_RDIV(fabs(wm[i][j]), ewt[j], ewi[j]);
      an = pw_max(an, sum * ewt[i]);
    }
    pdnorm = PW_DIV(an, fabs(hl0));
#pragma unroll
    for (int i = 0; i < N; ++i) wm[i][i] += 1.;
    // dgefa: Gaussian elimination with partial pivoting, multipliers stored negated.
    // The pivot row is a run-time index; rows are picked / exchanged with selects.
#pragma unroll
    for (int k = 0; k < N - 1; ++k) {
      int lp = k;
      double amax = fabs(wm[k][k]);
#pragma unroll
      for (int i = k + 1; i < N; ++i)
        if (fabs(wm[i][k]) > amax) { amax = fabs(wm[i][k]); lp = i; }
      ipvt[k] = lp;
      double piv = wm[k][k];
#pragma unroll
      for (int i = k + 1; i < N; ++i)
        if (i == lp) piv = wm[i][k];
      if (piv == 0.) { ierpj = 1; continue; }
#pragma unroll
      for (int i = k + 1; i < N; ++i)
        if (i == lp) wm[i][k] = wm[k][k];
      wm[k][k] = piv;
      const double t = PW_DIV(-1., piv);
#pragma unroll
      for (int i = k + 1; i < N; ++i) wm[i][k] *= t;
#pragma unroll
      for (int j = k + 1; j < N; ++j) {
        double tj = wm[k][j];
#pragma unroll
        for (int i = k + 1; i < N; ++i)
          if (i == lp) { tj = wm[i][j]; wm[i][j] = wm[k][j]; }
        wm[k][j] = tj;
#pragma unroll
        for (int i = k + 1; i < N; ++i) wm[i][j] += tj * wm[i][k];
      }
    }
    ipvt[N - 1] = N - 1;
    if (wm[N - 1][N - 1] == 0.) ierpj = 1;
#pragma unroll
    for (int k = 0; k < N; ++k) rdiag[k] = PW_DIV(1., wm[k][k]);
  }
  // solsy -> dgesl (job = 0)
  PW_HD void solsy(double* b) {
    iersl = 0;
#pragma unroll
    for (int k = 0; k < N - 1; ++k) {
      const int lp = ipvt[k];
      double t = b[k];
#pragma unroll
      for (int i = k + 1; i < N; ++i)
        if (i == lp) { t = b[i]; b[i] = b[k]; }
      b[k] = t;
#pragma unroll
      for (int i = k + 1; i < N; ++i) b[i] += t * wm[i][k];
    }
#pragma unroll
    for (int kb = 0; kb < N; ++kb) {
      const int k = N - 1 - kb;
      b[k] = PW_RDIV(b[k], wm[k][k], rdiag[k]);
      const double t = -b[k];
#pragma unroll
      for (int i = 0; i < k; ++i) b[i] += t * wm[i][k];
    }
  }

  // stoda corrector loop (labels 220-430).  Returns true when the corrector failed to
  // converge (label 430: the caller retracts, and retries with h/4 or gives up).
  PW_HD bool correction(double pnorm, double* del_out, int* m_out) {
    int m = 0;
    double rate = 0., del = 0., delp = 0.;
#pragma unroll
    for (int i = 0; i < N; ++i) y[i] = yh[1][i];
    f.at(tn);
    for (;;) {
      // the only evaluation of the right-hand side in the step loop (prja has the other)
      f.eval(y, savf);
      ++nfe;
      if (m == 0) {
        if (ipup > 0) {
          prja();
          ipup = 0;
          rc = 1.;
          nslp = nst;
          crate = 0.7;
          if (ierpj != 0) return true;
        }
#pragma unroll
        for (int i = 0; i < N; ++i) acor[i] = 0.;
      }
      if (miter == 0) {
#pragma unroll
        for (int i = 0; i < N; ++i) {
          savf[i] = h * savf[i] - yh[2][i];
          y[i] = savf[i] - acor[i];
        }
        del = vmnorm(y);
#pragma unroll
        for (int i = 0; i < N; ++i) {
          y[i] = yh[1][i] + el0 * savf[i];
          acor[i] = savf[i];
        }
      } else {
#pragma unroll
        for (int i = 0; i < N; ++i) y[i] = h * savf[i] - (yh[2][i] + acor[i]);
        solsy(y);
        del = vmnorm(y);
#pragma unroll
        for (int i = 0; i < N; ++i) {
          acor[i] += y[i];
          y[i] = yh[1][i] + el0 * acor[i];
        }
      }
      // convergence test (label 400)
      if (del <= 100. * pnorm * kEps) break;
      if (m != 0 || meth != 1) {
        if (m != 0) {
          double rm = 1024.;
          if (del <= (1024. * delp)) rm = PW_DIV(del, delp);
          rate = (rm > rate) ? rm : rate;
          const double c2 = 0.2 * crate;
          crate = (rm > c2) ? rm : c2;
        }
        const double c15 = 1.5 * crate;
        const double dcon = PW_RDIV(del * ((c15 < 1.) ? c15 : 1.), tq2 * conit, rtc);
        if (dcon <= 1.) {
          if (rate != 0.) {  // (rate = 0: pd = 0, or NaN for h = 0, neither of which changes pdest >= 0)
            const double pd = PW_DIV(rate, fabs(h * el0));
            pdest = (pd > pdest) ? pd : pdest;
          }
          if (pdest != 0.) pdlast = pdest;
          break;
        }
      }
      ++m;
      if (m == 3 /*maxcor*/ || (m >= 2 && del > 2. * delp)) {
        if (miter == 0 || jcur == 1) return true;
        ipup = miter;
        // restart the corrector with a fresh Jacobian (label 220)
        m = 0; rate = 0.; del = 0.;
#pragma unroll
        for (int i = 0; i < N; ++i) y[i] = yh[1][i];
      } else {
        delp = del;
      }
    }
    *del_out = del;
    *m_out = m;
    return false;
  }

  PW_HD void methodswitch(double dsm, double pnorm, double* pdh, double* rh) {
    if (meth == 1) {
      int nqm2;
      double rh2;
      if (nq > 5) return;
      if (dsm <= (100. * pnorm * kEps) || pdest == 0.) {
        if (irflag == 0) return;
        rh2 = 2.;
        nqm2 = nq < mxords ? nq : mxords;
      } else {
        double rh1 = pw_step_ratio(dsm, l, 1.2, 0.0000012);
        double rh1it = 2. * rh1;
        *pdh = pdlast * fabs(h);
        if ((*pdh * rh1) > 0.00001) rh1it = PW_DIV(sm1(nq), *pdh);
        rh1 = pw_min(rh1, rh1it);
        // (ODEPACK's branch for nq > mxords cannot be taken: nq <= 5 = mxords here)
        const double dm2 = dsm * PW_DIV(tab.cm1[nq], tab.cm2[nq]);
        rh2 = pw_step_ratio(dm2, l, 1.2, 0.0000012);
        nqm2 = nq;
        if (rh2 < ratio * rh1) return;
      }
      *rh = rh2;
      icount = 20;
      meth = 2;
      miter = jtyp;
      pdlast = 0.;
      nq = nqm2;
      l = nq + 1;
      return;
    }
    // currently BDF: consider Adams
    // Exact shortcut.  On a successful step dsm <= 1, hence rh2 >= 1/(1.2 + 1.2e-6) = 0.83333.
    // The stability bound gives rh1 <= sm1(nq)/pdh <= 0.575/pdh (or rh1 <= 1e-5/pdh when the
    // bound is not applied), so for pdh = pdnorm*|h| >= 1 the test `rh1*ratio < 5*rh2`
    // below is true whatever the two roots are: stay with BDF without evaluating them.
    // (In the stiff regime, where BDF is in use, pdh >> 1 on almost every step.)
    *pdh = pdnorm * fabs(h);
    if (ratio == 5. && *pdh >= 1.) return;
    // (ODEPACK's branch for mxordn < nq cannot be taken: nq <= 5 < 12 = mxordn here)
    double dm1 = dsm * PW_DIV(tab.cm2[nq], tab.cm1[nq]);
    double rh1 = pw_step_ratio(dm1, l, 1.2, 0.0000012);
    const int nqm1 = nq, lm1 = l;
    double rh1it = 2. * rh1;
    *pdh = pdnorm * fabs(h);
    if ((*pdh * rh1) > 0.00001) rh1it = PW_DIV(sm1(nqm1), *pdh);
    rh1 = pw_min(rh1, rh1it);
    const double rh2 = pw_step_ratio(dsm, l, 1.2, 0.0000012);
    if ((rh1 * ratio) < (5. * rh2)) return;
    const double alpha = pw_max(0.001, rh1);
    dm1 *= pw_root(alpha, lm1);
    if (dm1 <= 1000. * kEps * pnorm) return;
    *rh = rh1;
    icount = 20;
    meth = 1;
    miter = 0;
    pdlast = 0.;
    nq = nqm1;
    l = nq + 1;
  }

  // labels 540-630.  orderflag: 0 no change, 1 h changes, 2 h and nq change
  PW_HD void orderswitch(double rhup, double dsm, double* pdh, double* rh, int* orderflag) {
    *orderflag = 0;
    double rhsm = pw_step_ratio(dsm, l, 1.2, 0.0000012);
    double rhdn = 0.;
    if (nq != 1) {
      double top[N];
      row_get(l, top);
      const double ddn = PW_RDIV(vmnorm(top), tq1, rtq1);
      rhdn = pw_step_ratio(ddn, nq, 1.3, 0.0000013);
    }
    if (meth == 1) {
      *pdh = pw_max(fabs(h) * pdlast, 0.000001);
      if (l < lmax) rhup = pw_min(rhup, PW_DIV(sm1(l), *pdh));
      rhsm = pw_min(rhsm, PW_DIV(sm1(nq), *pdh));
      if (nq > 1) rhdn = pw_min(rhdn, PW_DIV(sm1(nq - 1), *pdh));
      pdest = 0.;
    }
    int newq;
    if (rhsm >= rhup) {
      if (rhsm >= rhdn) {
        newq = nq;
        *rh = rhsm;
      } else {
        newq = nq - 1;
        *rh = rhdn;
        if (kflag < 0 && *rh > 1.) *rh = 1.;
      }
    } else {
      if (rhup <= rhdn) {
        newq = nq - 1;
        *rh = rhdn;
        if (kflag < 0 && *rh > 1.) *rh = 1.;
      } else {
        *rh = rhup;
        if (*rh >= 1.1) {
          const double r = PW_DIV(elp[l], (double)l);
          nq = l;
          l = nq + 1;
          double top[N];
#pragma unroll
          for (int i = 0; i < N; ++i) top[i] = acor[i] * r;
          row_set(l, top);
          *orderflag = 2;
          return;
        }
        ialth = 3;
        return;
      }
    }
    if (meth == 1) {
      if ((*rh * *pdh * 1.00001) < sm1(newq))
        if (kflag == 0 && *rh < 1.1) { ialth = 3; return; }
    } else {
      if (kflag == 0 && *rh < 1.1) { ialth = 3; return; }
    }
    if (kflag <= -2) *rh = pw_min(*rh, 0.2);
    if (newq == nq) { *orderflag = 1; return; }
    nq = newq;
    l = nq + 1;
    *orderflag = 2;
  }

  PW_HD void endstoda() {  // labels 700-720
#pragma unroll
    for (int i = 0; i < N; ++i) acor[i] *= rtq2u;
    hold = h;
    jstart = 1;
  }

  // One step of the core integrator (DSTODA).  ODEPACK's goto web is laid out as ONE loop
  // with one copy of each block -- coefficient reload (label 150), rescale (175-180),
  // predict + correct (200-450), retract, order selection (540-630) -- selected by flags, so
  // that the step loop stays inside the instruction cache.  hmin = 0 throughout.
  PW_HD void stoda() {
    kflag = 0;
    const double told = tn;
    int ncf = 0;
    ierpj = 0; iersl = 0; jcur = 0;
    double rh = 0., pdh = 0., del = 0., dsm = 0.;
    int m = 0;
    bool reload = false;   // label 150 pending
    bool rescale = false;  // labels 175-180 pending ...
    bool finish = false;   // ... followed by rmax = 10 and label 700 (successful step)
    if (jstart == 0) {
      lmax = maxord + 1;
      nq = 1; l = 2; ialth = 2; rmax = 10000.; rc = 0.; el0 = 1.; crate = 0.7;
      hold = h; nslp = 0; ipup = miter;
      icount = 20; irflag = 0; pdest = 0.; pdlast = 0.; ratio = 5.;
      meth_tab = 1;
      reload = true;
    }
    if (jstart == -1) {
      ipup = miter;
      lmax = maxord + 1;
      if (ialth == 1) ialth = 2;
      if (meth != mused) {
        meth_tab = meth;  // call dcfode(meth, elco, tesco)
        ialth = l;
        reload = true;
      }
      if (h != hold) {
        rh = PW_DIV(h, hold);
        h = hold;
        rescale = true;
      }
    }
    for (;;) {
      if (reload) {
        resetcoeff();
        reload = false;
      }
      if (rescale) {
        scaleh(&rh, &pdh);
        rescale = false;
        if (finish) {
          rmax = 10.;
          endstoda();
          return;
        }
      }
      // prediction + corrector
      if (fabs(rc - 1.) > 0.3 /*ccmax*/) ipup = miter;
      if (nst >= nslp + 20 /*msbp*/) ipup = miter;
      tn += h;
      // predict (yh <- yh * Pascal), correct, test; after a failure the same call site -- one
      // copy of the code -- takes the prediction back (labels 430 / 500 start the same way)
      double pnorm = 0., sgn = 1.;
      bool cfail = false, efail = false;
#pragma unroll 1
      for (int undo = 0; undo < 2; ++undo) {
        pascal(sgn);
        if (undo) break;
        pnorm = vmnorm(yh[1]);
        cfail = correction(pnorm, &del, &m);
        if (!cfail) {
          jcur = 0;
          dsm = PW_RDIV((m == 0) ? del : vmnorm(acor), tq2, rtq2);
          efail = !(dsm <= 1.);
        }
        if (!(cfail || efail)) break;
        tn = told;
        rmax = 2.;
        sgn = -1.;
      }
      if (cfail) {
        ++ncf;
        if (h == 0. || ncf == 10 /*mxncf*/) {
          kflag = -2;
          hold = h;
          jstart = 1;
          return;
        }
        rh = 0.25;
        ipup = miter;
        rescale = true;
        continue;
      }
      double rhup = 0.;
      if (!efail) {
        // successful step
        kflag = 0;
        ++nst;
#if defined(PW_LSODA_HIST) && !defined(__CUDA_ARCH__)
        ++pw_lsoda_hist[meth - 1][nq];
#endif
        hu = h;
        nqu = nq;
        rtq2u = rtq2;
        mused = meth;
        // yh(., j) += el(j) * acor for j = 1..l: one guarded copy (el(j) = 0 in the table for j > l
        // would do as well, but rows above l must keep their bits)
#pragma unroll
        for (int j = 1; j <= kF + 1; ++j) {
          const bool on = (j <= l);
          const double r = elp[j];
#pragma unroll
          for (int i = 0; i < N; ++i)
            if (on) yh[j][i] += r * acor[i];
        }
        if (PW_UNLIKELY(nq > kF)) {
#pragma unroll 1
          for (int j = kF + 2; j <= l; ++j) {
            const double r = elp[j];
            for (int i = 0; i < N; ++i) yhs_[(j) * N + i] += r * acor[i];
          }
        }
        --icount;
        if (icount < 0) {
          methodswitch(dsm, pnorm, &pdh, &rh);
          if (meth != mused) {
            rescale = true;
            finish = true;
            continue;
          }
        }
        --ialth;
        if (ialth != 0) {
          if (!(ialth > 1 || l == lmax)) row_set(lmax, acor);
          endstoda();
          return;
        }
        if (l != lmax) {
          double top[N];
          row_get(lmax, top);
#pragma unroll
          for (int i = 0; i < N; ++i) savf[i] = acor[i] - top[i];
          const double dup = PW_RDIV(vmnorm(savf), tq3, rtq3);
          rhup = pw_step_ratio(dup, l + 1, 1.4, 0.0000014);
        }
      } else {
        // error test failed (label 500)
        --kflag;
        if (h == 0.) {
          kflag = -1; hold = h; jstart = 1;
          return;
        }
        if (kflag <= -3) {
          // three or more failures: reset to order 1 (label 640)
          if (kflag == -10) {
            kflag = -1; hold = h; jstart = 1;
            return;
          }
          rh = 0.1;
          h *= rh;
#pragma unroll
          for (int i = 0; i < N; ++i) y[i] = yh[1][i];
          f.at_cold(tn);
          f.eval_cold(y, savf);
          ++nfe;
#pragma unroll
          for (int i = 0; i < N; ++i) yh[2][i] = h * savf[i];
          ipup = miter;
          ialth = 5;
          if (nq != 1) {
            nq = 1;
            l = 2;
            reload = true;
          }
          continue;
        }
      }
      // order / step-size selection (labels 540-630), after a success with ialth = 0 or
      // after a first or second error-test failure
      int orderflag;
      orderswitch(rhup, dsm, &pdh, &rh, &orderflag);
      if (!efail) {
        if (orderflag == 0) {
          endstoda();
          return;
        }
        finish = true;
      } else if (orderflag == 0) {
        // label 610 reached from a failed step: "ialth = 3; go to 700" cannot
        // happen here because kflag < 0 skips the 10 percent test
        rh = (rh < 0.2) ? rh : 0.2;
      }
      if (orderflag == 2) reload = true;
      rescale = true;
    }
  }

  // DINTDY with k = 0
  PW_HD int intdy(double t, double* dky) const {
    const double tp = tn - hu - 100. * kEps * copysign(fabs(tn) + fabs(hu), hu);
    if ((t - tp) * (t - tn) > 0.) return -2;
    const double s = PW_DIV(t - tn, h);
#pragma unroll
    for (int i = 0; i < N; ++i) dky[i] = 0.;
    if (PW_UNLIKELY(nq > kF)) {
      for (int i = 0; i < N; ++i) dky[i] = yhs_[(l) * N + i];
#pragma unroll 1
      for (int j = l - 1; j >= kF + 2; --j)
        for (int i = 0; i < N; ++i) dky[i] = yhs_[(j) * N + i] + s * dky[i];
    }
    // Horner from row l down; rows above l are skipped, and row l itself enters as
    // yh[l] + s * 0 (dky is still zero there) -- one predicated fma per row and component
#pragma unroll
    for (int j = kF + 1; j >= 1; --j) {
      const bool on = (j <= l);
#pragma unroll
      for (int i = 0; i < N; ++i) {
        const double hn = pw_fma(s, dky[i], yh[j][i]);
        dky[i] = on ? hn : dky[i];
      }
    }
    return 0;
  }

  // ---- driver (DLSODA), split so that callers can run many integrations in lock step:
  //   begin()        first-call initialisation (istate = 1 block, initial step size)
  //   advance_one()  one pass of the main loop: per-step checks + one stoda() step
  //   reached(tout)  itask = 1 stopping test;  intdy(tout, y) output by interpolation
  //   new_call()     what a continuation call (istate = 2) does first: nslast = nst
  // call() strings them together exactly like one `lsoda(..., itask = 1)` invocation.
  PW_HD void reset() { started = false; }

  PW_HD int begin(double t0, const double* y0, double tout) {
    started = true;
    mxordn = kLsodaMaxOrdN; mxords = kLsodaMaxOrdS;
    hmxi = (o.hmax > 0.) ? 1. / o.hmax : 0.;
    tn = t0; tsw = t0; maxord = mxordn;
    jstart = 0; nhnil = 0; nst = 0; nje = 0; nslast = 0; hu = 0.; nqu = 0; mused = 0; miter = 0;
    meth = 1; jtyp = 2;
#pragma unroll
    for (int j = 1; j <= kF + 1; ++j)
#pragma unroll
      for (int i = 0; i < N; ++i) yh[j][i] = 0.;
    for (int j = 0; j <= kQ + 1; ++j)
      for (int i = 0; i < N; ++i) yhs_[(j) * N + i] = 0.;
#pragma unroll
    for (int i = 0; i < N; ++i) { y[i] = y0[i]; yh[1][i] = y0[i]; }
    f.at_cold(tn);
    f.eval_cold(y, yh[2]);
    nfe = 1;
    nq = 1;
    h = 1.;
    if (!ewset(yh[1])) return -3;
    double h0 = o.h0;
    if (!(h0 > 0.)) {
      const double tdist = fabs(tout - t0);
      const double w0 = pw_max(fabs(t0), fabs(tout));
      if (tdist < 2. * kEps * w0) return -3;
      double tol = o.rtol;
      if (tol <= 0.)
        for (int i = 0; i < N; ++i) {
          const double ayi = fabs(y[i]);
          if (ayi != 0.) tol = pw_max(tol, o.atol / ayi);
        }
      tol = pw_max(tol, 100. * kEps);
      tol = pw_min(tol, 0.001);
      double sum = vmnorm(yh[2]);
      sum = 1. / (tol * w0 * w0) + tol * sum * sum;
      h0 = 1. / sqrt(sum);
      h0 = pw_min(h0, tdist);
      h0 = copysign(h0, tout - t0);
    }
    const double rh = fabs(h0) * hmxi;
    if (rh > 1.) h0 /= rh;
    h = h0;
#pragma unroll
    for (int i = 0; i < N; ++i) yh[2][i] *= h0;
    return 0;
  }

  PW_HD void new_call() { nslast = nst; }
  PW_HD bool reached(double tout) const { return (tn - tout) * h >= 0.; }

  // Returns 0 after a successful step, < 0 (ODEPACK istate) on failure.
  PW_HD int advance_one() {
    if (nst != 0 || jstart != 0) {  // label 250 (skipped on the very first pass)
      if ((nst - nslast) >= o.mxstep) return -1;
      if (!ewset(yh[1])) return -6;
    }
    // "too much accuracy requested" (tolsf = eps * ||y / (rtol |y| + atol)||_max > 1) needs
    // rtol < eps: with atol >= 0 every weighted component is below 1 / rtol.  Not evaluated otherwise.
    if (!(o.rtol >= kEps && o.atol >= 0.)) {
      const double tolsf = kEps * vmnorm(yh[1]);
      if (tolsf > 1.) return (nst == 0) ? -3 : -2;
    }
    if ((tn + h) == tn) ++nhnil;
    stoda();
    if (kflag == 0) {
      if (meth != mused) {
        tsw = tn;
        maxord = mxordn;
        if (meth == 2) maxord = mxords;
        jstart = -1;
      }
      return 0;
    }
    return (kflag == -1) ? -4 : -5;
  }

  PW_HD void current(double* yout) const {
#pragma unroll
    for (int i = 0; i < N; ++i) yout[i] = yh[1][i];
  }

#if defined(PW_HE_PS_KERNEL)
  // ---- the same step, cut into phases (experiment, see he_population_ps in pw_helium.cuh) -------
  // advance_one() + stoda() as a resumable state machine: every ps_*() below runs ONE block of the
  // step -- exactly the statements of the code above, in the same order -- and returns the phase the
  // integration is in afterwards.  The helium kernel keeps one model per lane and lets the WARP decide
  // which block runs next (the phase most of its lanes are waiting in), so that lanes execute a block
  // together from its first instruction instead of drifting through the step at different offsets;
  // see he_population_ps (pw_helium.cuh).  The locals of stoda() that live across blocks are members.
  enum { kPsPredict = 0, kPsIter = 1, kPsJac = 2, kPsTest = 3, kPsFail = 4, kPsOrder = 5, kPsStepDone = 6 };
  double ps_told, ps_rh, ps_pdh, ps_del, ps_dsm, ps_pnorm, ps_rhup, ps_rate, ps_delp;
  int ps_ncf, ps_m, ps_rc;
  bool ps_reload, ps_rescale, ps_finish, ps_cfail, ps_efail, ps_fresh;

  PW_HD void ps_new_step() { ps_fresh = true; }   // the next ps_predict() starts a new advance_one()

  PW_HD int ps_step_done() {   // tail of advance_one()
    if (kflag == 0) {
      if (meth != mused) {
        tsw = tn;
        maxord = mxordn;
        if (meth == 2) maxord = mxords;
        jstart = -1;
      }
      ps_rc = 0;
    } else {
      ps_rc = (kflag == -1) ? -4 : -5;
    }
    return kPsStepDone;
  }

  PW_HD int ps_predict() {
    if (ps_fresh) {
      ps_fresh = false;
      if (nst != 0 || jstart != 0) {  // label 250 (skipped on the very first pass)
        if ((nst - nslast) >= o.mxstep) { ps_rc = -1; return kPsStepDone; }
        if (!ewset(yh[1])) { ps_rc = -6; return kPsStepDone; }
      }
      if (!(o.rtol >= kEps && o.atol >= 0.)) {
        const double tolsf = kEps * vmnorm(yh[1]);
        if (tolsf > 1.) { ps_rc = (nst == 0) ? -3 : -2; return kPsStepDone; }
      }
      if ((tn + h) == tn) ++nhnil;
      // stoda() entry
      kflag = 0;
      ps_told = tn;
      ps_ncf = 0;
      ps_cfail = false; ps_efail = false;
      ierpj = 0; iersl = 0; jcur = 0;
      ps_rh = 0.; ps_pdh = 0.; ps_del = 0.; ps_dsm = 0.;
      ps_m = 0;
      ps_reload = false; ps_rescale = false; ps_finish = false;
      if (jstart == 0) {
        lmax = maxord + 1;
        nq = 1; l = 2; ialth = 2; rmax = 10000.; rc = 0.; el0 = 1.; crate = 0.7;
        hold = h; nslp = 0; ipup = miter;
        icount = 20; irflag = 0; pdest = 0.; pdlast = 0.; ratio = 5.;
        meth_tab = 1;
        ps_reload = true;
      }
      if (jstart == -1) {
        ipup = miter;
        lmax = maxord + 1;
        if (ialth == 1) ialth = 2;
        if (meth != mused) {
          meth_tab = meth;
          ialth = l;
          ps_reload = true;
        }
        if (h != hold) {
          ps_rh = PW_DIV(h, hold);
          h = hold;
          ps_rescale = true;
        }
      }
    }
    if (ps_reload) {
      resetcoeff();
      ps_reload = false;
    }
    if (ps_rescale) {
      scaleh(&ps_rh, &ps_pdh);
      ps_rescale = false;
      if (ps_finish) {
        rmax = 10.;
        endstoda();
        return ps_step_done();
      }
    }
    if (fabs(rc - 1.) > 0.3 /*ccmax*/) ipup = miter;
    if (nst >= nslp + 20 /*msbp*/) ipup = miter;
    tn += h;
    pascal(1.);
    ps_pnorm = vmnorm(yh[1]);
    // correction(): labels 220 ff.
    ps_m = 0; ps_rate = 0.; ps_del = 0.; ps_delp = 0.;
#pragma unroll
    for (int i = 0; i < N; ++i) y[i] = yh[1][i];
    f.at(tn);
    return kPsIter;
  }

  // second half of a corrector iteration: chord / functional update, convergence test
  PW_HD int ps_iter_b() {
    double del;
    if (ps_m == 0) {
#pragma unroll
      for (int i = 0; i < N; ++i) acor[i] = 0.;
    }
    if (miter == 0) {
#pragma unroll
      for (int i = 0; i < N; ++i) {
        savf[i] = h * savf[i] - yh[2][i];
        y[i] = savf[i] - acor[i];
      }
      del = vmnorm(y);
#pragma unroll
      for (int i = 0; i < N; ++i) {
        y[i] = yh[1][i] + el0 * savf[i];
        acor[i] = savf[i];
      }
    } else {
#pragma unroll
      for (int i = 0; i < N; ++i) y[i] = h * savf[i] - (yh[2][i] + acor[i]);
      solsy(y);
      del = vmnorm(y);
#pragma unroll
      for (int i = 0; i < N; ++i) {
        acor[i] += y[i];
        y[i] = yh[1][i] + el0 * acor[i];
      }
    }
    ps_del = del;
    bool conv = (del <= 100. * ps_pnorm * kEps);
    if (!conv && (ps_m != 0 || meth != 1)) {
      if (ps_m != 0) {
        double rm = 1024.;
        if (del <= (1024. * ps_delp)) rm = PW_DIV(del, ps_delp);
        ps_rate = (rm > ps_rate) ? rm : ps_rate;
        const double c2 = 0.2 * crate;
        crate = (rm > c2) ? rm : c2;
      }
      const double c15 = 1.5 * crate;
      const double dcon = PW_RDIV(del * ((c15 < 1.) ? c15 : 1.), tq2 * conit, rtc);
      if (dcon <= 1.) {
        if (ps_rate != 0.) {
          const double pd = PW_DIV(ps_rate, fabs(h * el0));
          pdest = (pd > pdest) ? pd : pdest;
        }
        if (pdest != 0.) pdlast = pdest;
        conv = true;
      }
    }
    if (conv) return kPsTest;
    ++ps_m;
    if (ps_m == 3 /*maxcor*/ || (ps_m >= 2 && del > 2. * ps_delp)) {
      if (miter == 0 || jcur == 1) { ps_cfail = true; return kPsFail; }
      ipup = miter;
      ps_m = 0; ps_rate = 0.; ps_del = 0.;
#pragma unroll
      for (int i = 0; i < N; ++i) y[i] = yh[1][i];
    } else {
      ps_delp = del;
    }
    return kPsIter;
  }

  PW_HD int ps_iter() {
    f.eval(y, savf);
    ++nfe;
    if (ps_m == 0 && ipup > 0) return kPsJac;   // this iteration needs the Jacobian first
    return ps_iter_b();
  }

  PW_HD int ps_jac() {
    prja();
    ipup = 0;
    rc = 1.;
    nslp = nst;
    crate = 0.7;
    if (ierpj != 0) { ps_cfail = true; return kPsFail; }
    return ps_iter_b();
  }

  PW_HD int ps_test() {
    jcur = 0;
    ps_dsm = PW_RDIV((ps_m == 0) ? ps_del : vmnorm(acor), tq2, rtq2);
    if (!(ps_dsm <= 1.)) { ps_cfail = false; ps_efail = true; return kPsFail; }
    ps_efail = false;
    // successful step
    kflag = 0;
    ++nst;
    hu = h;
    nqu = nq;
    rtq2u = rtq2;
    mused = meth;
#pragma unroll
    for (int j = 1; j <= kF + 1; ++j) {
      const bool on = (j <= l);
      const double r = elp[j];
#pragma unroll
      for (int i = 0; i < N; ++i)
        if (on) yh[j][i] += r * acor[i];
    }
    if (PW_UNLIKELY(nq > kF)) {
#pragma unroll 1
      for (int j = kF + 2; j <= l; ++j) {
        const double r = elp[j];
        for (int i = 0; i < N; ++i) yhs_[(j) * N + i] += r * acor[i];
      }
    }
    --icount;
    if (icount < 0) {
      methodswitch(ps_dsm, ps_pnorm, &ps_pdh, &ps_rh);
      if (meth != mused) {
        ps_rescale = true;
        ps_finish = true;
        return kPsPredict;
      }
    }
    --ialth;
    if (ialth != 0) {
      if (!(ialth > 1 || l == lmax)) row_set(lmax, acor);
      endstoda();
      return ps_step_done();
    }
    ps_rhup = 0.;
    if (l != lmax) {
      double top[N];
      row_get(lmax, top);
#pragma unroll
      for (int i = 0; i < N; ++i) savf[i] = acor[i] - top[i];
      const double dup = PW_RDIV(vmnorm(savf), tq3, rtq3);
      ps_rhup = pw_step_ratio(dup, l + 1, 1.4, 0.0000014);
    }
    return kPsOrder;
  }

  // corrector failure (label 430) or error-test failure (label 500): take the prediction back
  PW_HD int ps_fail() {
    tn = ps_told;
    rmax = 2.;
    pascal(-1.);
    if (ps_cfail) {
      ps_cfail = false;
      ++ps_ncf;
      if (h == 0. || ps_ncf == 10 /*mxncf*/) {
        kflag = -2;
        hold = h;
        jstart = 1;
        return ps_step_done();
      }
      ps_rh = 0.25;
      ipup = miter;
      ps_rescale = true;
      return kPsPredict;
    }
    --kflag;
    if (h == 0.) {
      kflag = -1; hold = h; jstart = 1;
      return ps_step_done();
    }
    if (kflag <= -3) {
      if (kflag == -10) {
        kflag = -1; hold = h; jstart = 1;
        return ps_step_done();
      }
      ps_rh = 0.1;
      h *= ps_rh;
#pragma unroll
      for (int i = 0; i < N; ++i) y[i] = yh[1][i];
      f.at_cold(tn);
      f.eval_cold(y, savf);
      ++nfe;
#pragma unroll
      for (int i = 0; i < N; ++i) yh[2][i] = h * savf[i];
      ipup = miter;
      ialth = 5;
      if (nq != 1) {
        nq = 1;
        l = 2;
        ps_reload = true;
      }
      return kPsPredict;
    }
    ps_rhup = 0.;
    return kPsOrder;
  }

  PW_HD int ps_order() {
    int orderflag;
    orderswitch(ps_rhup, ps_dsm, &ps_pdh, &ps_rh, &orderflag);
    if (!ps_efail) {
      if (orderflag == 0) {
        endstoda();
        return ps_step_done();
      }
      ps_finish = true;
    } else if (orderflag == 0) {
      ps_rh = (ps_rh < 0.2) ? ps_rh : 0.2;
    }
    if (orderflag == 2) ps_reload = true;
    ps_rescale = true;
    return kPsPredict;
  }
#endif  // PW_HE_PS_KERNEL

  // One `lsoda(..., itask = 1)` call: advance to tout and interpolate.
  // First call initialises from (t0, y0).  Returns istate.
  PW_HD int call(double t0, const double* y0, double tout, double* yout) {
    if (!started) {
      const int rc = begin(t0, y0, tout);
      if (rc < 0) return rc;
    } else {
      new_call();
      if (reached(tout)) return (intdy(tout, yout) != 0) ? -3 : 2;
    }
    for (;;) {
      const int rc = advance_one();
      if (rc < 0) {
        current(yout);
        return rc;
      }
      if (!reached(tout)) continue;
      return (intdy(tout, yout) != 0) ? -3 : 2;
    }
  }
};

}  // namespace pw
