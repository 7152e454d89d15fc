// fp64 instantiation of the batched Cassie2d kernels (parity build: the legacy one-env ABI
// computes in double like the reference, and tests compare it with the oracle to 1e-9).
#include "launch.cuh"
namespace cassie {
template struct Launch<double>;
}
