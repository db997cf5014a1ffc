// Drives the handler body of ntk_b200_ffi.cc through the stand-in header (tests/ffi_mock/xla/ffi/api/ffi.h):
// builds the AnyBuffers XLA would pass and calls GramImpl on the given CUDA stream.
#include "../../neural-tangents_b200/csrc/ntk_b200_ffi.cc"

#ifndef NTK_B200_HAVE_XLA_FFI
#error "the stand-in header was not found: compile with -Itests/ffi_mock"
#endif

#include <cstring>

extern "C" int ntk_ffi_mock_gram(void* stream, int64_t prog, int64_t ctx, int32_t flags, int32_t is_f64, void* x1,
                                 const int64_t* d1, int32_t nd1, void* x2, const int64_t* d2, int32_t nd2,
                                 void* nngp, void* ntk, char* err, int32_t err_len) {
  const ffi::DataType t = is_f64 ? ffi::F64 : ffi::F32;
  ffi::AnyBuffer b1(t, x1, std::vector<int64_t>(d1, d1 + nd1)), b2(t, x2, std::vector<int64_t>(d2, d2 + nd2));
  const std::vector<int64_t> od = {d1[0], d2[0]};
  ffi::Error e = GramImpl(static_cast<cudaStream_t>(stream), prog, ctx, flags, b1, b2,
                          ffi::Result<ffi::AnyBuffer>(ffi::AnyBuffer(t, nngp, od)),
                          ffi::Result<ffi::AnyBuffer>(ffi::AnyBuffer(t, ntk, od)));
  if (e.failure() && err && err_len > 0) {
    strncpy(err, e.message().c_str(), err_len - 1);
    err[err_len - 1] = 0;
  }
  return e.success() ? 0 : 1;
}
