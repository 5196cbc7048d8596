// density_grid.cpp -- host half of PathTracer::AddDensityDataToVolume after the OpenVDB file read (PT/PathTracer.cpp:1391-1452):
// MaxDensityInTheGrid, the AABB the volume takes, the 32 x 32 x 32 grid of majorants for empty-space skipping, and the temperature patch.
// The reference then converts the FloatGrid to NanoVDB and uploads that; this framework uploads the dense values themselves (engine.cu), which is
// what a NanoVDB read at an index coordinate inside the root bbox returns.  No CUDA in here: b200pt_prepare_density_grid runs without a GPU.
#include "host_api.h"
#include <algorithm>
#include <cmath>
#include <cstring>

namespace b200pt {

void prepare_density_grid(const b200pt_density_grid &g, PreparedGrid &out) {
    if (!g.Density) throw CudaError{ B200PT_ERR_WRONG_ARGUMENTS, "density grid: Density is NULL" };
    for (int k = 0; k < 3; k++) if (g.Dim[k] == 0u || g.Dim[k] > 4096u) throw CudaError{ B200PT_ERR_WRONG_ARGUMENTS, "density grid: Dim must be 1..4096 per axis" };
    if (!(g.VoxelSize > 0.0)) throw CudaError{ B200PT_ERR_WRONG_ARGUMENTS, "density grid: VoxelSize must be positive" };
    const int nx = (int)g.Dim[0], ny = (int)g.Dim[1], nz = (int)g.Dim[2];
    const size_t count = (size_t)nx * ny * nz;
    out.values.assign(g.Density, g.Density + count);
    out.max_density = *std::max_element(out.values.begin(), out.values.end());       // tools::minMax(tree, true).max(), :1393
    if (!(out.max_density > 0.0f)) throw CudaError{ B200PT_ERR_WRONG_ARGUMENTS, "density grid: no positive density value" };
    float tmin = g.TemperatureMin, tmax = g.TemperatureMax;
    if (g.Temperature && !(tmin < tmax)) { const auto mm = std::minmax_element(g.Temperature, g.Temperature + count); tmin = *mm.first; tmax = *mm.second; }
    out.has_temperature = g.Temperature != nullptr;
    // :1408-1420 the index bbox, scaled so that its largest |coordinate| becomes 1
    float largest = 0.0f;
    for (int k = 0; k < 3; k++) {
        out.imin[k] = g.IndexMin[k]; out.dim[k] = (int)g.Dim[k];
        out.corner_min[k] = (float)g.IndexMin[k]; out.corner_max[k] = (float)(g.IndexMin[k] + (int)g.Dim[k] - 1);
        largest = std::max(largest, std::max(std::fabs(out.corner_min[k]), std::fabs(out.corner_max[k])));
    }
    for (int k = 0; k < 3; k++) { out.corner_min[k] /= largest; out.corner_max[k] /= largest; }
    // :1422-1452 one pass over the bbox, Y walked from the top (the image-space flip of :1432)
    out.max_densities.assign(32768, 0.0f);
    for (int z = 0; z < nz; z++) {
        const int cz = (z * 32) / nz;
        for (int y = 0; y < ny; y++) {
            const int cy = (y * 32) / ny;
            float *row = out.values.data() + ((size_t)z * ny + (size_t)(ny - 1 - y)) * nx;
            const float *trow = g.Temperature ? g.Temperature + ((size_t)z * ny + (size_t)(ny - 1 - y)) * nx : nullptr;
            for (int x = 0; x < nx; x++) {
                const float normalised = std::min(std::max(row[x] / out.max_density, 0.0f), 1.0f);
                float &cell = out.max_densities[(size_t)((x * 32) / nx) + (size_t)cy * 32 + (size_t)cz * 1024];
                if (cell < normalised) cell = normalised;
                if (trow) {                                                           // the normalised temperature is written INTO the density grid (:1447-1449)
                    const float t = std::max((trow[x] - tmin) / (tmax - tmin), 0.0f);
                    if (t > 0.0f) row[x] = t;
                }
            }
        }
    }
    // NanoVDB's grid header as SampleNanoVDBBuffer reads it (SH/Volume.slang:73-87,99): world bbox = the index bbox [min, max + 1] through the map
    // (GridStats), rounded outwards to integers; the map's float inverse scale and translation
    for (int k = 0; k < 3; k++) {
        const double wlo = (double)g.IndexMin[k] * g.VoxelSize + g.Translation[k], whi = (double)(g.IndexMin[k] + (int)g.Dim[k]) * g.VoxelSize + g.Translation[k];
        const int lo = (int)std::floor((float)wlo), hi = (int)std::ceil((float)whi);
        out.wmin[k] = (float)lo; out.wext[k] = (float)(hi - lo);
        out.inv_vs[k] = (float)(1.0 / g.VoxelSize); out.trans[k] = (float)g.Translation[k];
    }
}

} // namespace b200pt
