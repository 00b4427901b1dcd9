// The beam-search instantiations that also track pyctcdecode's word frames (text_frames, what HF
// turns into word_offsets: HF:models/wav2vec2_with_lm/processing_wav2vec2_with_lm.py:416-443). A
// translation unit of its own so that it compiles next to beam.cu instead of after it.
#include "beam_launch.cuh"

namespace coral {

int32_t launch_beam_frames(coral_decoder* dec, BeamLaunch& L, int32_t B, int32_t beam_width, cudaStream_t st) {
  if (beam_width <= 32) return launch_beam<32, 32, 128, true>(dec, L, B, st);
  if (beam_width <= 64) return launch_beam<64, 64, 192, true>(dec, L, B, st);
  if (beam_width <= 128) return launch_beam<128, 128, 320, true>(dec, L, B, st);
  if (beam_width <= 256) return launch_beam<256, 256, 640, true>(dec, L, B, st);
  return launch_beam<256, 512, 1280, true>(dec, L, B, st);
}

}  // namespace coral
