// Fused octave kernel (placeholder until the fused path lands: reports "not
// supported" so that every octave runs on the generic per-stage kernels).
#include "common.cuh"

namespace sb {

  bool fused_octave_supported(const Taps*, int)
  {
    return false;
  }

  void launch_fused_octave(const OctaveDesc&, const OctaveDesc*, int, const Taps*, int, cudaStream_t)
  {
  }

}  // namespace sb
