// ml_tma_flat.cu -- the TMA-staged steric kernels for grids whose rows are not a multiple of 16 bytes.
//
// A strided tensor map cannot describe such a field (global strides must be multiples of 16 bytes), but a rank-1 map
// over the flat array can, and a box may start at any element: the stage of a level is filled by one 1-D box per row
// (k_steric_tma<..., FLAT = true>::refill_stage in ml_tma.cu).  Same kernels, same arithmetic, their own translation
// unit: see the note above launch_segment_flat in ml_tma.cu.
#define ML_TMA_FLAT_TU 1
#include "ml_tma.cu"
