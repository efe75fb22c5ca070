// stands in for <cuda_bf16.h> when the kernel sources are compiled for the host (tests/simt_emu)
#include "../cuda_emu.h"
