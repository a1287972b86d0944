// Host harness for mogp_b200/csrc/wspace.cuh: the SAME phase functions the CUDA kernels call, run under sequential
// loops (a phase has no dependencies between its work items).  stdin: m, mean_func, n_theta, then x[m], y[m], then n_theta
// rows (a, variance, lengthscale, noise); stdout: per theta "lml dA dvar dls dnoise info rho".
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include "../../mogp_b200/csrc/wspace.cuh"

using namespace mogp;

int main() {
  int m = 0, mean_func = 1, nth = 0;
  if (scanf("%d %d %d", &m, &mean_func, &nth) != 3 || m <= 0) return 2;
  std::vector<double> x(m), y(m);
  for (int i = 0; i < m; ++i)
    if (scanf("%lf", &x[i]) != 1) return 2;
  for (int i = 0; i < m; ++i)
    if (scanf("%lf", &y[i]) != 1) return 2;
  std::vector<double> ws(WS_WORK_ELEMS, 0.0), C((size_t)m * WS_R);
  WsWork w{ws.data()};
  WsNodes nd;
  ws_make_nodes(*std::min_element(x.begin(), x.end()), *std::max_element(x.begin(), x.end()), &nd);
  for (int j = 0; j < WS_R; ++j) w.t()[j] = nd.t[j];
  for (int k = 0; k < m; ++k) ws_basis_row(nd, x[k], C.data() + (size_t)k * WS_R);                 // k_ws_basis
  for (int e = 0; e < WS_GRAM_ENTRIES; ++e) ws_gram_entry(C.data(), x.data(), y.data(), m, e, w);  // k_ws_gram
  for (int q = 0; q < nth; ++q) {                                                                   // k_ws_eval, phase by phase
    WsTheta th;
    th.mean_func = mean_func;
    if (scanf("%lf %lf %lf %lf", &th.a, &th.var, &th.ls, &th.noise) != 4) return 2;
    w.out()[6] = (double)WS_R;
    for (int e = 0; e < WS_RR; ++e) ws_ph_kernel(w, th, e);
    for (int k = 0; k < WS_R; ++k) {
      ws_pc_pivot(w, th, k);
      if (w.sc()[7] == 0.0) break;
      for (int j = 0; j < WS_R; ++j) ws_pc_row(w, k, j);
    }
    const int rho = (int)w.out()[6];
    for (int e = 0; e < rho * WS_R; ++e) ws_ph_rs(w, e);
    for (int e = 0; e < rho * rho; ++e) ws_ph_a(w, th, e / rho, e % rho);
    ws_ph_chol(w, w.A(), rho);
    if (w.out()[5] == 0.0) {
      for (int j = 0; j < WS_R; ++j) ws_ph_x(w, w.A(), rho, j);
      for (int e = 0; e < WS_RR; ++e) ws_ph_w(w, th, rho, e);
      ws_ph_finish(w, th, w.A(), rho);
    }
    printf("%.17g %.17g %.17g %.17g %.17g %d %d\n", w.out()[0], w.out()[1], w.out()[2], w.out()[3], w.out()[4], (int)w.out()[5], rho);
  }
  return 0;
}
