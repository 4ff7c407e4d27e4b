// spb_strips.cpp -- see spb_strips.h
#include "spb_strips.h"

#include <algorithm>
#include <limits>

namespace spb {

void partition_rows(uint32_t height, uint32_t quantum, uint32_t parts, const double *rowCost, uint32_t *bounds)
{
    if (parts == 0) return;
    if (quantum == 0) quantum = 1;
    const uint32_t rows = (height + quantum - 1) / quantum;
    auto to_pixels = [&](uint32_t r) { return std::min<uint64_t>((uint64_t)r * quantum, height); };
    std::vector<uint32_t> cuts(parts + 1, rows);
    cuts[0] = 0;
    if (rows <= parts || !rowCost)
    {
        // even split in whole quantum rows (earlier parts take the remainder); one row each when
        // there are no more rows than parts
        const uint32_t base = rows / parts, extra = rows % parts;
        uint32_t at = 0;
        for (uint32_t p = 0; p < parts; ++p)
        {
            cuts[p] = at;
            at += base + (p < extra ? 1u : 0u);
        }
        cuts[parts] = rows;
    }
    else
    {
        // cost floor: a row of zero cost still has to go somewhere; keeps the optimum unique enough
        double top = 0.0;
        for (uint32_t r = 0; r < rows; ++r) top = std::max(top, rowCost[r]);
        const double floorCost = 1e-9 * std::max(1.0, top);
        std::vector<double> prefix(rows + 1, 0.0);
        for (uint32_t r = 0; r < rows; ++r) prefix[r + 1] = prefix[r] + std::max(rowCost[r], floorCost);
        const double inf = std::numeric_limits<double>::infinity();
        // best[k][i]: smallest possible largest sum when the first i rows form k non-empty strips
        std::vector<double> best((size_t)(parts + 1) * (rows + 1), inf);
        std::vector<uint32_t> arg((size_t)(parts + 1) * (rows + 1), 0);
        auto at = [&](uint32_t k, uint32_t i) -> size_t { return (size_t)k * (rows + 1) + i; };
        best[at(0, 0)] = 0.0;
        for (uint32_t k = 1; k <= parts; ++k)
            for (uint32_t i = k; i + (parts - k) <= rows; ++i)
            {
                double b = inf;
                uint32_t bj = k - 1;
                for (uint32_t j = k - 1; j < i; ++j)
                {
                    const double v = std::max(best[at(k - 1, j)], prefix[i] - prefix[j]);
                    if (v < b) { b = v; bj = j; }
                }
                best[at(k, i)] = b;
                arg[at(k, i)] = bj;
            }
        uint32_t i = rows;
        for (uint32_t k = parts; k >= 1; --k)
        {
            cuts[k] = i;
            i = arg[at(k, i)];
        }
        cuts[0] = 0;
    }
    for (uint32_t p = 0; p <= parts; ++p) bounds[p] = (uint32_t)to_pixels(cuts[p]);
}

std::vector<double> row_seconds(uint32_t height, uint32_t quantum, uint32_t parts, const uint32_t *bounds,
                                const double *units, const double *seconds)
{
    if (quantum == 0) quantum = 1;
    const uint32_t rows = (height + quantum - 1) / quantum;
    std::vector<double> out(rows, 0.0);
    for (uint32_t p = 0; p < parts; ++p)
    {
        const uint32_t b = bounds[p], e = bounds[p + 1];
        if (e <= b) continue;
        const uint32_t r0 = b / quantum, r1 = (e + quantum - 1) / quantum;
        double total = 0.0;
        for (uint32_t r = r0; r < r1 && r < rows; ++r) total += units[r];
        const double scale = total > 0.0 ? seconds[p] / total : 0.0;
        for (uint32_t r = r0; r < r1 && r < rows; ++r) out[r] = units[r] * scale;
    }
    return out;
}

} // namespace spb
