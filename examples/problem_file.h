/** examples/problem_file.h — reader of the binary problem dump written by ShellProblem.save (gsstructuralanalysis_b200/problem.py):
    'KLP1', 16 int32 header, 8 doubles, the arrays in the order of kl_problem, then (optional) the Neumann sides. */
#pragma once
#include <cstdint>
#include <fstream>
#include <string>
#include <vector>

#include "../include/kl_shell.h"

struct ProblemFile {
    kl_problem P{};
    std::vector<double> U1, U2, cp, w, fixed, pl_uv, pl_val, neu_val;
    std::vector<int32_t> map, neu_side;
    bool load(const char* path) {
        std::ifstream f(path, std::ios::binary);
        char magic[4];
        int32_t h[16];
        double d[8];
        if (!f.read(magic, 4) || std::string(magic, 4) != "KLP1") return false;
        f.read((char*)h, sizeof(h));
        f.read((char*)d, sizeof(d));
        const int ncp = h[4], has_w = h[5], npl = h[15];
        auto rd = [&](std::vector<double>& v, size_t n) { v.resize(n); f.read((char*)v.data(), sizeof(double) * n); };
        rd(U1, h[2]); rd(U2, h[3]); rd(cp, 3 * (size_t)ncp);
        if (has_w) rd(w, ncp);
        map.resize(3 * (size_t)ncp);
        f.read((char*)map.data(), sizeof(int32_t) * map.size());
        rd(fixed, h[7]);
        if (npl) { rd(pl_uv, 2 * (size_t)npl); rd(pl_val, 3 * (size_t)npl); }
        if (!f) return false;
        int32_t nneu = 0;
        if (f.read((char*)&nneu, sizeof(nneu)) && nneu > 0) {
            neu_side.resize(nneu);
            f.read((char*)neu_side.data(), sizeof(int32_t) * nneu);
            rd(neu_val, 3 * (size_t)nneu);
            if (!f) return false;
        } else {
            nneu = 0;
        }
        P.degree[0] = h[0]; P.degree[1] = h[1]; P.n_knots[0] = h[2]; P.n_knots[1] = h[3];
        P.knots[0] = U1.data(); P.knots[1] = U2.data(); P.cp = cp.data(); P.weights = has_w ? w.data() : nullptr;
        P.dof_map = map.data(); P.n_free = h[6]; P.n_fixed = h[7]; P.fixed_values = h[7] ? fixed.data() : nullptr;
        P.material = h[8]; P.compressible = h[9]; P.num_gauss_thickness = h[10]; P.bending = h[11]; P.metric_z2 = h[12];
        P.quA = h[13]; P.quB = h[14]; P.n_point_loads = npl;
        P.E = d[0]; P.nu = d[1]; P.thickness = d[2]; P.mr_ratio = d[3];
        P.body_force[0] = d[4]; P.body_force[1] = d[5]; P.body_force[2] = d[6]; P.pressure = d[7];
        P.point_load_uv = npl ? pl_uv.data() : nullptr; P.point_load_val = npl ? pl_val.data() : nullptr;
        P.n_neumann = nneu; P.neumann_side = nneu ? neu_side.data() : nullptr; P.neumann_val = nneu ? neu_val.data() : nullptr;
        return true;
    }
};
