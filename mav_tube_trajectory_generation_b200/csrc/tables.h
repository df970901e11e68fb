#ifndef MTG_TABLES_H_
#define MTG_TABLES_H_

#define MTG_TAB_LD 12   // Polynomial::kMaxN (reference polynomial.h:45)
#define MTG_BASE_LD 22  // Polynomial::kMaxConvolutionSize (reference polynomial.h:48)

namespace mtg {

struct Tables {
  int N;
  int derivative;
  double H1[MTG_TAB_LD * MTG_TAB_LD];     // row-major, leading dimension MTG_TAB_LD
  double Ainv1[MTG_TAB_LD * MTG_TAB_LD];
  // rank factor of H1: H1 = W^T W, W is (N-d) x N. W = L^T Ainv1[d..N-1, :] with
  // Q(1)[d.., d..] = L L^T. The cost is evaluated as a sum of squares |W dhat|^2,
  // which has no cancellation across terms (a direct dhat^T H1 dhat loses ~6 digits).
  double W[MTG_TAB_LD * MTG_TAB_LD];
  // Lt(i, a) = L[a][i] for a >= i (upper triangular, (N-d) x (N-d)): |W dhat|^2 = |Lt chat|^2 with
  // chat_j = c_j T^j the scaled coefficients, j = d..N-1 (the kernels have chat at hand)
  double Lt[MTG_TAB_LD * MTG_TAB_LD];
  double base[MTG_BASE_LD * MTG_BASE_LD]; // base_coefficients_: B(n,i) = i!/(i-n)!
};

// false on invalid (N, derivative)
bool compute_tables(int N, int derivative, Tables* out);

}  // namespace mtg
#endif
