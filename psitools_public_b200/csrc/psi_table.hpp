// psi_table.hpp -- host-side construction of the level-major node table and of DistParams from the
// C-ABI descriptor.  Shared by the CUDA library and by the test-only lane-emulation harness.
#pragma once
#include <string.h>
#include <vector>

#include "../../include/psi_b200.h"
#include "psi_quad.cuh"

namespace psi {

// Reorder TanhSinh.xj/wj (tanhsinh.py:65-66, original order) level-major: level 0 = multiples of
// 2^max_m, level m = odd multiples of 2^(max_m-m).  off has max_m + 2 entries.
inline bool build_level_major(const double* xj, const double* wj, int n_nodes, int max_level,
                              std::vector<node_t>& lv, std::vector<int>& off) {
    lv.clear();
    lv.reserve(n_nodes);
    off.assign(max_level + 2, 0);
    for (long long j = 0; j < n_nodes; j += (1ll << max_level)) lv.push_back(node_t{xj[j], wj[j]});
    off[1] = (int)lv.size();
    for (int m = 1; m <= max_level; ++m) {
        const long long st = 1ll << (max_level - m);
        for (long long j = st; j < n_nodes; j += 2 * st) lv.push_back(node_t{xj[j], wj[j]});
        off[m + 1] = (int)lv.size();
    }
    return (int)lv.size() == n_nodes;
}

inline DistParams dist_from_desc(const psi_dist_desc& s) {
    DistParams d;
    memset(&d, 0, sizeof(d));
    d.kind = s.kind; d.single_size = s.single_size;
    d.mu = s.mu; d.taumin = s.tau_min; d.taumax = s.tau_max; d.c = s.sound_speed;
    d.rhod = s.rhod; d.q = s.q;
    d.beta = s.beta; d.norm_pl = s.norm_pl; d.amp = s.amp; d.curv = s.curv;
    d.aL = s.aL; d.aP = s.aP; d.aBT = s.aBT; d.scale = s.scale;
    d.n_init_poles = d.n_calc_poles = s.n_poles;
    for (int i = 0; i < s.n_poles; ++i) {
        d.init_poles[i] = s.poles[i];                  // psi_dispersion.py:67 (unscaled)
        d.calc_poles[i] = s.poles[i] / s.tau_max;      // psi_dispersion.py:410
    }
    return d;
}

}  // namespace psi
