// XLA-FFI handler over the C-ABI of include/ntk_b200.h: the `jax.ffi` binding north_star asks for.
//
// One handler, `NtkGram`: kernel_fn(x1, x2, ('nngp', 'ntk')) on raw inputs, i.e. the body of `kernel_fn_x1`
// (`_src/stax/requirements.py:939-953`).  x1, x2 and the two results are XLA device buffers and the work is enqueued
// on XLA's own compute stream (`ffi::PlatformStream<cudaStream_t>` -> `ntk_gram_device_on_stream`): nothing is copied,
// nothing synchronises, the handler returns as soon as the kernels are in the stream -- the contract of an XLA custom
// call.  `prog` / `ctx` are the handles returned by ntk_program_create / ntk_context_create (one context per device, as
// PjRt runs one host thread per device under pmap: `_src/batching.py:775`), passed as integer attributes.
//
// jaxlib's header (`xla/ffi/api/ffi.h`, from `jax.ffi.include_dir()`) is not in this image:
//   make ffi JAX_INCLUDE=$(python -c "import jax.ffi; print(jax.ffi.include_dir())")     # -> libntk_b200_ffi.so
// Without it the file compiles to an empty object.  tests/test_ffi_shim.py compiles this very file against a
// stand-in for the API subset used below (tests/ffi_mock/) and runs the handler body on the GPU.
// INTEGRATION.md §2 shows the Python side and embeds this file verbatim (checked by the same test).
#if defined(__has_include)
#if __has_include("xla/ffi/api/ffi.h")
#define NTK_B200_HAVE_XLA_FFI 1
#endif
#endif

#ifdef NTK_B200_HAVE_XLA_FFI
#include <cuda_runtime_api.h>

#include <cstdint>

#include "ntk_b200.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

static ffi::Error GramImpl(cudaStream_t stream, int64_t prog, int64_t ctx, int32_t flags, ffi::AnyBuffer x1,
                           ffi::AnyBuffer x2, ffi::Result<ffi::AnyBuffer> nngp, ffi::Result<ffi::AnyBuffer> ntk) {
  const auto d = x1.dimensions();                       // [n1, H, W, C] or [n1, C]
  const bool img = d.size() == 4;
  if (!img && d.size() != 2)
    return ffi::Error(ffi::ErrorCode::kInvalidArgument, "x1 must be [n, H, W, C] or [n, C]");
  if (x1.element_type() != ffi::F32 && x1.element_type() != ffi::F64)
    return ffi::Error(ffi::ErrorCode::kInvalidArgument, "x1 must be float32 or float64");
  const int dtype = x1.element_type() == ffi::F64 ? NTK_F64 : NTK_F32;
  const int n1 = static_cast<int>(d[0]), n2 = static_cast<int>(x2.dimensions()[0]);
  const bool want_ntk = (flags & NTK_FLAG_NTK) != 0;
  const int rc = ntk_gram_device_on_stream(
      reinterpret_cast<ntk_context_t*>(ctx), reinterpret_cast<const ntk_program_t*>(prog), dtype, x1.untyped_data(), n1,
      x2.untyped_data(), n2, img ? static_cast<int>(d[1]) : 0, img ? static_cast<int>(d[2]) : 0,
      static_cast<int>(d.back()), static_cast<uint32_t>(flags), nngp->untyped_data(),
      want_ntk ? ntk->untyped_data() : nullptr, /*ld=*/n2, nullptr, nullptr, stream);
  return rc == NTK_OK ? ffi::Error::Success() : ffi::Error(ffi::ErrorCode::kInternal, ntk_last_error());
}

XLA_FFI_DEFINE_HANDLER_SYMBOL(NtkGram, GramImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<int64_t>("prog")
                                  .Attr<int64_t>("ctx")
                                  .Attr<int32_t>("flags")
                                  .Arg<ffi::AnyBuffer>()
                                  .Arg<ffi::AnyBuffer>()
                                  .Ret<ffi::AnyBuffer>()
                                  .Ret<ffi::AnyBuffer>());
#endif  // NTK_B200_HAVE_XLA_FFI
