// Host-side MultiKtensor (reference src/multi_ktensor.cpp): see include/multi_ktensor.h for why it exists next to the
// device-side queue.
#include <algorithm>

#include "multi_ktensor.h"
#include "utils/utils.h"

namespace cals {

MultiKtensor::MultiKtensor(vector<dim_t> &modes, dim_t buffer_size)
    : Ktensor(buffer_size, modes), tensor_modes(modes), owner(buffer_size, 0) {
  for (Matrix &f : get_factors()) {
    f.zero();
    f.resize(f.get_rows(), 0);
  }
}

int MultiKtensor::first_fit(dim_t width) const {
  dim_t run = 0;
  for (dim_t c = 0; c < owner.size(); c++) {
    run = owner[c] == 0 ? run + 1 : 0;
    if (run == width)
      return static_cast<int>(c + 1 - width);
  }
  throw BufferFull();
}

void MultiKtensor::refresh_views() {
  dim_t active = owner.size();
  while (active > 0 && owner[active - 1] == 0)
    active--;
  start = 0;
  for (Matrix &f : get_factors()) {
    f.reset_data();
    f.resize(f.get_rows(), active);
  }
}

MultiKtensor &MultiKtensor::add(Ktensor &ktensor) {
  const dim_t R = ktensor.get_components();
  const int col = first_fit(R);
  vector<double *> where(ktensor.get_n_modes());
  dim_t n = 0;
  for (Matrix &f : get_factors())
    where[n++] = f.reset_data().get_data() + static_cast<dim_t>(col) * f.get_rows();
  ktensor.attach(where);

  const dim_t id = next_id++;
  std::fill(owner.begin() + col, owner.begin() + col + R, id);
  vector<Matrix> gramians;
  for (const Matrix &f : ktensor.get_factors()) {
    gramians.emplace_back(R, R);
    ops::update_gramian(f, gramians.back());
  }
  ktensor.set_iters(1);
  flag_jk = flag_jk || ktensor.is_jk();

  RegistryEntry entry{ktensor, std::move(gramians), col, id};
  if (line_search) {
    entry.ls_params.prev_ktensor = Ktensor(R, tensor_modes);
    entry.ls_params.backup_ktensor = Ktensor(R, tensor_modes);
    entry.ls_params.cuda = cuda;
    entry.ls_params.interval = ls_params.interval;
    entry.ls_params.step = ls_params.step;
    entry.ls_params.method = ls_params.method;
    entry.ls_params.T = ls_params.T;
  }
  registry.insert(std::pair<int, RegistryEntry>(static_cast<int>(id), std::move(entry)));
  refresh_views();
  return *this;
}

MultiKtensor &MultiKtensor::remove(dim_t ktensor_id) {
  RegistryEntry &entry = registry.at(static_cast<int>(ktensor_id));
  entry.ktensor.detach();
  std::replace(owner.begin(), owner.end(), ktensor_id, dim_t(0));
  registry.erase(static_cast<int>(ktensor_id));
  refresh_views();
  return *this;
}

MultiKtensor &MultiKtensor::compress() {
  // models in column order; each moves left by the number of free columns before it
  dim_t holes = 0;
  for (dim_t c = 0; c < owner.size();) {
    if (owner[c] == 0) {
      holes++;
      c++;
      continue;
    }
    const dim_t id = owner[c];
    RegistryEntry &entry = registry.at(static_cast<int>(id));
    const dim_t R = entry.ktensor.get_components();
    if (holes) {
      vector<double *> where(get_n_modes());
      dim_t n = 0;
      for (Matrix &f : entry.ktensor.get_factors()) {
        double *old = f.get_data();
        where[n] = old - holes * f.get_rows();
        std::copy(old, old + f.get_n_elements(), where[n]); // leftwards: forward copy is safe when ranges overlap
        std::fill(std::max(old, where[n] + f.get_n_elements()), old + f.get_n_elements(), 0.0);
        f.attach(where[n]);
        n++;
      }
      std::fill(owner.begin() + c - holes, owner.begin() + c - holes + R, id);
      std::fill(owner.begin() + std::max(c, c - holes + R), owner.begin() + c + R, dim_t(0));
      entry.col -= static_cast<int>(holes);
    }
    c += R;
  }
  refresh_views();
  return *this;
}

} // namespace cals
