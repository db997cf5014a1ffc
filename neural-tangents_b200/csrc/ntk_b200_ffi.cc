// XLA-FFI handler over the C-ABI of include/ntk_b200.h: the `jax.ffi` binding north_star asks for.
//
// jaxlib's headers (`xla/ffi/api/ffi.h`, from `jax.ffi.include_dir()`) are not in this image, so this
// file compiles to an empty object here; with them present, `make ffi JAX_INCLUDE=$(python -c "import jax.ffi;
// print(jax.ffi.include_dir())")` builds libntk_b200_ffi.so, and INTEGRATION.md §2 shows the Python side
// (`jax.ffi.register_ffi_target('ntk_b200_gram', jax.ffi.pycapsule(lib.NtkGram), platform='CUDA')`).
//
// The handler is a thin wrapper: x1, x2 and the two result buffers are XLA device buffers, nothing is copied;
// `prog` / `ctx` are the handles returned by ntk_program_create / ntk_context_create, passed as integer attributes.
#if defined(__has_include)
#if __has_include("xla/ffi/api/ffi.h")
#define NTK_B200_HAVE_XLA_FFI 1
#endif
#endif

#ifdef NTK_B200_HAVE_XLA_FFI
#include <cstdint>

#include "../../include/ntk_b200.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

// kernel_fn(x1, x2, ('nngp', 'ntk')) on raw inputs (`_src/stax/requirements.py:939-953`).
static ffi::Error GramImpl(int64_t prog, int64_t ctx, int32_t flags, ffi::AnyBuffer x1, ffi::AnyBuffer x2,
                           ffi::Result<ffi::AnyBuffer> nngp, ffi::Result<ffi::AnyBuffer> ntk) {
  const auto d = x1.dimensions();                       // [n1, H, W, C] or [n1, C]
  const bool img = d.size() == 4;
  if (!img && d.size() != 2) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "x1 must be [n, H, W, C] or [n, C]");
  const int dtype = x1.element_type() == ffi::F64 ? NTK_F64 : NTK_F32;
  const int n2 = static_cast<int>(x2.dimensions()[0]);
  // The context owns the stream the work is enqueued on (one context per device, as PjRt runs one host thread
  // per device under pmap: `_src/batching.py:775`); XLA's stream is synchronised with it by the caller's
  // ntk_context_synchronize after the call when results are consumed by other XLA ops.
  const int rc = ntk_gram_device(reinterpret_cast<ntk_context_t*>(ctx), reinterpret_cast<const ntk_program_t*>(prog),
                                 dtype, x1.untyped_data(), static_cast<int>(d[0]), x2.untyped_data(), n2,
                                 img ? static_cast<int>(d[1]) : 0, img ? static_cast<int>(d[2]) : 0,
                                 static_cast<int>(d.back()), static_cast<uint32_t>(flags), nngp->untyped_data(),
                                 (flags & NTK_FLAG_NTK) ? ntk->untyped_data() : nullptr, n2, nullptr, nullptr);
  if (rc != NTK_OK) return ffi::Error(ffi::ErrorCode::kInternal, ntk_last_error());
  const int rs = ntk_context_synchronize(reinterpret_cast<ntk_context_t*>(ctx));
  return rs == NTK_OK ? ffi::Error::Success() : ffi::Error(ffi::ErrorCode::kInternal, ntk_last_error());
}

XLA_FFI_DEFINE_HANDLER_SYMBOL(NtkGram, GramImpl,
                              ffi::Ffi::Bind()
                                  .Attr<int64_t>("prog")
                                  .Attr<int64_t>("ctx")
                                  .Attr<int32_t>("flags")
                                  .Arg<ffi::AnyBuffer>()
                                  .Arg<ffi::AnyBuffer>()
                                  .Ret<ffi::AnyBuffer>()
                                  .Ret<ffi::AnyBuffer>());
#endif  // NTK_B200_HAVE_XLA_FFI
