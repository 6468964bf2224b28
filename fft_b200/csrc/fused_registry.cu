// fused_registry.cu -- collects the specialised kernels of every translation unit.
#include <cstdlib>

#include "fused.cuh"
#include "cluster.cuh"
#include "tiled.cuh"
#include "flat.cuh"
namespace ssfft {
void register_fused_f32_a(std::vector<FusedEntry> &);
void register_fused_f32_b(std::vector<FusedEntry> &);
void register_fused_f32_c(std::vector<FusedEntry> &);
void register_fused_f64_a(std::vector<FusedEntry> &);
void register_fused_f64_b(std::vector<FusedEntry> &);
void register_fused_f32_d(std::vector<FusedEntry> &);
void register_fused_f64_c(std::vector<FusedEntry> &);
void register_fused_f32_e(std::vector<FusedEntry> &);

const std::vector<FusedEntry> &fused_registry() {
    static const std::vector<FusedEntry> reg = [] {
        std::vector<FusedEntry> v;
        // SSFFT_DISABLE_FUSED=1 forces every size through the generic kernel (used by the parity tests
        // to exercise both paths); it never selects a CPU path -- there is none.
        const char *off = getenv("SSFFT_DISABLE_FUSED");
        if (off && off[0] == '1') return v;
        register_fused_f32_a(v);
        register_fused_f32_b(v);
        register_fused_f32_c(v);
        register_fused_f64_a(v);
        register_fused_f64_b(v);
        register_fused_f32_d(v);
        register_fused_f64_c(v);
        register_fused_f32_e(v);
        return v;
    }();
    return reg;
}

void register_tile_f32_a(std::vector<TileEntry> &);
void register_tile_f32_b(std::vector<TileEntry> &);
void register_tile_f64_a(std::vector<TileEntry> &);

const std::vector<TileEntry> &tile_registry() {
    static const std::vector<TileEntry> reg = [] {
        std::vector<TileEntry> v;
        const char *off = getenv("SSFFT_DISABLE_TILED");  // parity tests: force the generic four-step
        if (off && off[0] == '1') return v;
        register_tile_f32_a(v);
        register_tile_f32_b(v);
        register_tile_f64_a(v);
        return v;
    }();
    return reg;
}

void register_fourstep_f32_a(std::vector<FourStepEntry> &);
void register_fourstep_f32_b(std::vector<FourStepEntry> &);
void register_fourstep_f32_c(std::vector<FourStepEntry> &);
void register_fourstep_f32_d(std::vector<FourStepEntry> &);
void register_fourstep_f64_a(std::vector<FourStepEntry> &);
void register_fourstep_f64_b(std::vector<FourStepEntry> &);

const std::vector<FourStepEntry> &fourstep_registry() {
    static const std::vector<FourStepEntry> reg = [] {
        std::vector<FourStepEntry> v;
        const char *off = getenv("SSFFT_DISABLE_CLUSTER");  // fall back to two launches per chunk
        if (off && off[0] == '1') return v;
        register_fourstep_f32_a(v);
        register_fourstep_f32_b(v);
        register_fourstep_f32_c(v);
        register_fourstep_f32_d(v);
        register_fourstep_f64_a(v);
        register_fourstep_f64_b(v);
        return v;
    }();
    return reg;
}
int fourstep_cluster_size() {
    static const int c = [] {
        const char *e = getenv("SSFFT_CLUSTER");
        int v = e ? atoi(e) : 4;
        return (v < 1 || v > 16) ? 4 : v;
    }();
    return c;
}

void register_flat_f32_a(std::vector<FlatEntry> &);
void register_flat_f32_b(std::vector<FlatEntry> &);
void register_flat_f32_c(std::vector<FlatEntry> &);
void register_flat_f32_d(std::vector<FlatEntry> &);
void register_flat_f32_e(std::vector<FlatEntry> &);
void register_flat_f32_h(std::vector<FlatEntry> &);
void register_flat_f32_i(std::vector<FlatEntry> &);
void register_flat_f32_j(std::vector<FlatEntry> &);
void register_flat_f32_k(std::vector<FlatEntry> &);
void register_flat_f64_a(std::vector<FlatEntry> &);

const std::vector<FlatEntry> &flat_registry() {
    static const std::vector<FlatEntry> reg = [] {
        std::vector<FlatEntry> v;
        const char *off = getenv("SSFFT_DISABLE_FLAT");  // fall back to the cluster kernels of tiled.cuh / cluster.cuh
        const char *off2 = getenv("SSFFT_DISABLE_TILED");
        if ((off && off[0] == '1') || (off2 && off2[0] == '1')) return v;
        register_flat_f32_a(v);
        register_flat_f32_b(v);
        register_flat_f32_c(v);
        register_flat_f32_d(v);
        register_flat_f32_e(v);
        register_flat_f32_h(v);
        register_flat_f32_i(v);
        register_flat_f32_j(v);
        register_flat_f32_k(v);
        register_flat_f64_a(v);
        return v;
    }();
    return reg;
}

void register_cluster_f32_a(std::vector<ClusterEntry> &);
void register_cluster_f32_b(std::vector<ClusterEntry> &);

const std::vector<ClusterEntry> &cluster_registry() {
    static const std::vector<ClusterEntry> reg = [] {
        std::vector<ClusterEntry> v;
        // SSFFT_DISABLE_DSMEM=1: fall back to the L2-scratch four-step (parity tests exercise both)
        const char *off = getenv("SSFFT_DISABLE_DSMEM");
        const char *off2 = getenv("SSFFT_DISABLE_TILED");
        if ((off && off[0] == '1') || (off2 && off2[0] == '1')) return v;
        register_cluster_f32_a(v);
        register_cluster_f32_b(v);
        return v;
    }();
    return reg;
}

// how many resident "waves" of CTAs a fused launch may create before CTAs start looping
int fused_waves() {
    static const int w = [] {
        const char *e = getenv("SSFFT_FUSED_WAVES");
        int v = e ? atoi(e) : 4;
        return v < 0 ? 0 : v;
    }();
    return w;
}
}  // namespace ssfft
