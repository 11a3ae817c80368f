// cals::Matrix -- column-major FP64 matrix on top of cals::Tensor (reference include/matrix.h:7-152,
// src/matrix.cpp:7-48).  col_stride == rows always, as in the reference.
#ifndef CALS_B200_MATRIX_H
#define CALS_B200_MATRIX_H

#include <cmath>

#include "tensor.h"

namespace cals {

class Matrix : public Tensor {
  dim_t rows{0};
  dim_t cols{0};
  dim_t col_stride{0};

public:
  Matrix() = default;
  ~Matrix() = default;
  Matrix(dim_t n_rows, dim_t n_cols) : Tensor{n_rows, n_cols}, rows{n_rows}, cols{n_cols}, col_stride{n_rows} {}
  Matrix(dim_t n_rows, dim_t n_cols, double *view_data)
      : Tensor{n_rows, n_cols, view_data}, rows{n_rows}, cols{n_cols}, col_stride{n_rows} {}

  Matrix(Matrix &&) = default;
  Matrix &operator=(Matrix &&) = default;
  Matrix(const Matrix &) = default;
  Matrix &operator=(const Matrix &) = default;

  [[nodiscard]] dim_t get_rows() const noexcept { return rows; }
  [[nodiscard]] dim_t get_cols() const noexcept { return cols; }
  [[nodiscard]] dim_t get_col_stride() const noexcept { return col_stride; }

  double &operator()(dim_t r, dim_t c) { return get_data()[r + c * col_stride]; }
  double operator()(dim_t r, dim_t c) const noexcept { return get_data()[r + c * col_stride]; }

  // Change the logical shape inside the allocation (reference include/matrix.h:69-76).
  Matrix &resize(dim_t new_rows, dim_t new_cols) noexcept {
    vector<dim_t> shape = {new_rows, new_cols};
    Tensor::resize(new_rows * new_cols, shape);
    rows = new_rows;
    cols = new_cols;
    col_stride = new_rows;
    return *this;
  }

  // Element-wise product with a matrix of the same size (reference src/matrix.cpp:12-17).
  Matrix &hadamard(const Matrix &other) {
    double *d = get_data();
    for (dim_t i = 0; i < get_n_elements(); i++)
      d[i] *= other[i];
    return *this;
  }

  void attach(double *where) { set_data(where); } // point at foreign storage (a multi-buffer column block)
  void detach() { reset_data(); }                 // back to the owned storage

  void print(const std::string &&text = "Matrix") const;
  void info() const;

  // max over columns of the sum of magnitudes (reference include/matrix.h:113-121)
  [[nodiscard]] double one_norm() const {
    double best = -DBL_MAX;
    for (dim_t c = 0; c < cols; c++) {
      double s = 0.0;
      for (dim_t r = 0; r < rows; r++)
        s += std::fabs((*this)(r, c));
      best = std::max(best, s);
    }
    return best;
  }

  // this (rows x cols, row-major image) <- transpose of rhs (reference src/matrix.cpp:40-47)
  Matrix &transpose_copy(const Matrix &rhs) {
    for (dim_t i = 0; i < rows; i++)
      for (dim_t j = 0; j < cols; j++)
        get_data()[j + i * cols] = rhs.get_data()[i + j * rows];
    return *this;
  }
};

} // namespace cals
#endif
