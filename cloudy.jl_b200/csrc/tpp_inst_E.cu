// explicit instantiations of the thread-per-parcel kernel, group E (split for parallel compilation)
#include <type_traits>
#include "tpp_kernel.cuh"
#include "tpp_instances.inc"
namespace cloudy {
tpp_fn tpp_lookup_E(int N, int P, int model) {
#define X(NN, PP)                                                                                          \
    if (N == NN && P == PP)                                                                                \
        return model == MODEL_RAINSHAFT ? (tpp_fn)tpp_kernel<NN, PP, MODEL_RAINSHAFT>                     \
               : model == MODEL_BOX_MOVING ? (tpp_fn)tpp_kernel<NN, PP, MODEL_BOX_MOVING> : (tpp_fn)tpp_kernel<NN, PP, MODEL_BOX>;
    TPP_SHAPES_E
#undef X
    return nullptr;
}
}  // namespace cloudy
