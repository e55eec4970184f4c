// Self-test of the race check: a 32-thread "kernel" in which every lane reads its neighbour's shared-memory write, with
// and without the __syncwarp that makes that legal.  tests/test_kernels_racecheck.py expects ThreadSanitizer to report
// the second and not the first.
#include "cuda_emul.h"

float probe_buf[32], probe_out[32];   // external linkage: the accesses must survive optimisation

template <bool SYNC>
static void probe_kernel(int) {
    const unsigned lane = threadIdx.x & 31;
    __syncwarp();                                  // every lane is running from here on
    for (int it = 0; it < 64; ++it) {
        probe_buf[lane] = (float)(lane + it);
        if (SYNC) __syncwarp();
        probe_out[lane] = probe_buf[(lane + 1) & 31];
        if (SYNC) __syncwarp();
    }
}

extern "C" int probe(int sync) {
    if (sync) cuda_emul::launch(probe_kernel<true>, 1, 32, probe_buf, 0, 0);
    else cuda_emul::launch(probe_kernel<false>, 1, 32, probe_buf, 0, 0);
    return 0;
}
