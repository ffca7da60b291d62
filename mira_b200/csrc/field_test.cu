// field_test.cu — element-wise field kernels behind mira_test_field_op (unit-test hook).
#include "ctx.hpp"
#include "testgen.cuh"
namespace mira_host {
int test_field_op_dev(int field, int op, const void* a, const void* b, size_t n, void* out) {
  unsigned blocks = (unsigned)((n + 127) / 128);
  if (field == MIRA_FQ) mira::k_test_field<mira::FqTag><<<blocks, 128>>>(op, a, b, n, out);
  else mira::k_test_field<mira::FrTag><<<blocks, 128>>>(op, a, b, n, out);
  CU(cudaGetLastError());
  CU(cudaDeviceSynchronize());
  return MIRA_OK;
}
}  // namespace mira_host
