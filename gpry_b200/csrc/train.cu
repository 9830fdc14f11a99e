// placeholder -- replaced below
#include "state.cuh"
namespace gpry {
void factorize_device(gpry_state*, int, int, int, const double*, const double*, const double*,
                      const double*, double*, double*, double*, double*, int*, bool) {
  throw GpryError{GPRY_ERR_ARG, "factorize: not built yet"};
}
void lml_batched_device(gpry_state*, int, int, int, const double*, const double*, const double*,
                        const double*, int, double*, double*, int*) {
  throw GpryError{GPRY_ERR_ARG, "lml: not built yet"};
}
}  // namespace gpry
