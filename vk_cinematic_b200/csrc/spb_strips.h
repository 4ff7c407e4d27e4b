// spb_strips.h -- strips of image rows across the GPUs of one box: host logic.
//
// The reference's only parallelism is 64x64 tiles popped from an atomic queue by 16 threads over a
// shared read-only scene (reference src/main.cpp:731-759,819-844, src/tile.h:11-42).  ComputeTiles
// orders tiles row-major, so a contiguous range of pixel rows is a contiguous byte range of the
// RGBA-f32 image: every GPU renders one strip and the strips are gathered -- the one exchange
// step of the path.  Strip boundaries fall on multiples of `quantum` rows and are re-cut between
// frames from the measured cost of every quantum row.
#pragma once

#include <stdint.h>
#include <vector>

namespace spb {

// Cuts `height` rows into `parts` contiguous strips whose boundaries are multiples of `quantum`
// (the last one ends at `height`).  rowCost: cost of every quantum row (ceil(height / quantum)
// values) or null for an even split.  With costs the cut minimises the LARGEST strip sum -- the
// frame takes as long as the slowest GPU -- by dynamic programming over (strips used, rows
// covered); every strip keeps at least one quantum row while rows last, trailing strips are empty
// when there are fewer quantum rows than parts.  bounds: parts + 1 ascending row numbers.
void partition_rows(uint32_t height, uint32_t quantum, uint32_t parts, const double *rowCost, uint32_t *bounds);

// Per-quantum-row cost in seconds from what every part measured last frame: units[r] = cost units
// of quantum row r (sp_b200_RenderRows tileRowCost, whichever part rendered it), seconds[p] = the
// kernel time of part p, bounds = the cut those were measured with.  A row's cost is its units
// times its part's seconds per unit: the model's error (a unit of sky is not a unit of bunny)
// cancels inside a part and the next cut moves work from slow parts to fast ones.
std::vector<double> row_seconds(uint32_t height, uint32_t quantum, uint32_t parts, const uint32_t *bounds,
                                const double *units, const double *seconds);

} // namespace spb
