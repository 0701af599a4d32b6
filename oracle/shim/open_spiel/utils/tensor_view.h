// TEST INFRASTRUCTURE ONLY (oracle/): see ../spiel.h.
// Row-major rank-R view over a caller-owned float span, as the reference uses
// it in ObservationTensor (twixt.cc:79-80, 115-116): optional zero fill, then
// element access by an index tuple.
#ifndef ORACLE_SHIM_OPEN_SPIEL_UTILS_TENSOR_VIEW_H_
#define ORACLE_SHIM_OPEN_SPIEL_UTILS_TENSOR_VIEW_H_

#include <array>
#include <cstddef>

#include "open_spiel/spiel.h"

namespace open_spiel {

template <int Rank>
class TensorView {
 public:
  TensorView(absl::Span<float> values, const std::array<int, Rank>& shape,
             bool reset)
      : values_(values), shape_(shape) {
    std::size_t want = 1;
    for (int d = 0; d < Rank; ++d) want *= static_cast<std::size_t>(shape_[d]);
    if (want != values_.size())
      SpielFatalError("TensorView: span size does not match shape");
    if (reset)
      for (std::size_t i = 0; i < want; ++i) values_[i] = 0.0f;
  }

  float& operator[](const std::array<int, Rank>& idx) {
    std::size_t flat = 0;
    for (int d = 0; d < Rank; ++d)
      flat = flat * static_cast<std::size_t>(shape_[d]) +
             static_cast<std::size_t>(idx[d]);
    return values_[flat];
  }

 private:
  absl::Span<float> values_;
  std::array<int, Rank> shape_;
};

}  // namespace open_spiel

#endif  // ORACLE_SHIM_OPEN_SPIEL_UTILS_TENSOR_VIEW_H_
