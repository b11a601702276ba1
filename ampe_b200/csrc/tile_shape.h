// Tile shapes and block sizes of the fused kernels (swept on B200, profiles/README.md)
#pragma once
#ifndef AMPE_T2Y
#define AMPE_T2Y 16
#define AMPE_NT2 256
#endif
#ifndef AMPE_T3Y
#define AMPE_T3Y 4
#define AMPE_T3Z 4
#define AMPE_NT3 512
#endif
// 3D plane-marching kernel: column 32 x AMPE_MY, AMPE_MZ planes per block
#ifndef AMPE_MY
#define AMPE_MY 8
#define AMPE_MZ 32
#endif
// 2D persistent TMA kernel: tile 32 x AMPE_TMA_TY, AMPE_TMA_NT threads
#ifndef AMPE_TMA_TY
#define AMPE_TMA_TY 32
#define AMPE_TMA_NT 512
#endif
// resident blocks per SM the register allocation is capped for
#ifndef AMPE_MARCH_MINB
#define AMPE_MARCH_MINB 2
#endif
// split 3D launches (AMPE_B200_SPLIT3D): phase + quaternion part / composition part
#ifndef AMPE_SPLIT_MINB1
#define AMPE_SPLIT_MINB1 3
#endif
#ifndef AMPE_SPLIT_MINB2
#define AMPE_SPLIT_MINB2 3
#endif
// column height (rows of 32 cells per block) of the two split kernels
#ifndef AMPE_SPLIT_MY1
#define AMPE_SPLIT_MY1 AMPE_MY
#endif
#ifndef AMPE_SPLIT_MY2
#define AMPE_SPLIT_MY2 AMPE_MY
#endif
#ifndef AMPE_KKS_MINB
#define AMPE_KKS_MINB 4
#endif
