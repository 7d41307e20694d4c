// Build shim, NOT reference code: lets the reference's src/numint/lebedev.cpp compile without
// Eigen. That file only ever touches its matrix parameter through .data() and operator()(i,j)
// on a row-major (N x 4) double matrix, so this 20-line stand-in is sufficient.
// Used only by oracle/Makefile to build oracle/_ref/libocc_ref.so (test infrastructure).
#pragma once
#include <cstddef>
#include <vector>
namespace Eigen {
constexpr int Dynamic = -1;
constexpr int RowMajor = 1;
template <typename T, int R, int C, int O> class Matrix {
public:
  Matrix(int rows, int cols) : m_rows(rows), m_cols(cols), m_data(static_cast<size_t>(rows > 0 ? rows : 1) * cols) {}
  T *data() { return m_data.data(); }
  const T *data() const { return m_data.data(); }
  T &operator()(int i, int j) { return m_data[static_cast<size_t>(i) * m_cols + j]; }
  int rows() const { return m_rows; }
  int cols() const { return m_cols; }
private:
  int m_rows, m_cols;
  std::vector<T> m_data;
};
} // namespace Eigen
namespace occ::numint::grid {
Eigen::Matrix<double, Eigen::Dynamic, 4, Eigen::RowMajor> lebedev(int num_points);
int nearest_grid_level_at_or_above(int num_points);
int nearest_grid_level_below(int num_points);
} // namespace occ::numint::grid
