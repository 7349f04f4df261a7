// hostmath.cpp — TEST-ONLY host build of fusionsense_b200/csrc/fs_math.cuh.
// Lets the CPU test-suite check the per-Gaussian projection / SH formulas (forward and backward) that the
// sm_100a kernels inline, against fp64 autograd of the oracle, without a GPU.  Never linked into libfsb200.
#include "../../fusionsense_b200/csrc/fs_math.cuh"
#include <cstring>

extern "C" {

// cam: 12 floats of the view matrix rows 0..2, then fx, fy, cx, cy
void hm_project_fwd(int n, const float* cam16, const float* means, const float* quats, const float* scales,
                    int width, int height, float eps2d, float near_plane, float far_plane, float radius_clip,
                    int* radii, float* means2d, float* depths, float* conics, float* comps) {
    fs::Camera cam;
    memcpy(cam.V, cam16, 12 * sizeof(float));
    cam.fx = cam16[12]; cam.fy = cam16[13]; cam.cx = cam16[14]; cam.cy = cam16[15];
    for (int i = 0; i < n; ++i) {
        fs::ProjFwd o = fs::project_fwd(cam, means[3 * i], means[3 * i + 1], means[3 * i + 2], quats[4 * i],
                                        quats[4 * i + 1], quats[4 * i + 2], quats[4 * i + 3], scales[3 * i],
                                        scales[3 * i + 1], scales[3 * i + 2], width, height, eps2d, near_plane,
                                        far_plane, radius_clip);
        radii[i] = o.radius;
        means2d[2 * i] = o.mx; means2d[2 * i + 1] = o.my;
        depths[i] = o.depth;
        conics[3 * i] = o.ca; conics[3 * i + 1] = o.cb; conics[3 * i + 2] = o.cc;
        comps[i] = o.comp;
    }
}

void hm_project_bwd(int n, const float* cam16, const float* means, const float* quats, const float* scales,
                    int width, int height, float eps2d, const float* v_means2d, const float* v_depths,
                    const float* v_conics, const float* v_comps, float* v_means, float* v_quats, float* v_scales,
                    float* v_R /*[n,9]*/, float* v_t /*[n,3]*/) {
    fs::Camera cam;
    memcpy(cam.V, cam16, 12 * sizeof(float));
    cam.fx = cam16[12]; cam.fy = cam16[13]; cam.cx = cam16[14]; cam.cy = cam16[15];
    for (int i = 0; i < n; ++i) {
        fs::project_bwd(cam, means[3 * i], means[3 * i + 1], means[3 * i + 2], quats[4 * i], quats[4 * i + 1],
                        quats[4 * i + 2], quats[4 * i + 3], scales[3 * i], scales[3 * i + 1], scales[3 * i + 2],
                        width, height, eps2d, v_means2d[2 * i], v_means2d[2 * i + 1], v_depths[i], v_conics[3 * i],
                        v_conics[3 * i + 1], v_conics[3 * i + 2], v_comps[i], v_means + 3 * i, v_quats + 4 * i,
                        v_scales + 3 * i, v_R + 9 * i, v_t + 3 * i);
    }
}

void hm_sh_basis(int n, int degree, const float* dirs_unit, float* basis /*[n,16]*/, float* dx, float* dy,
                 float* dz) {
    for (int i = 0; i < n; ++i) {
        float b[16] = {0};
        fs::sh_basis(degree, dirs_unit[3 * i], dirs_unit[3 * i + 1], dirs_unit[3 * i + 2], b);
        memcpy(basis + 16 * i, b, sizeof(b));
        fs::sh_basis_grad(degree, dirs_unit[3 * i], dirs_unit[3 * i + 1], dirs_unit[3 * i + 2], dx + 16 * i,
                          dy + 16 * i, dz + 16 * i);
    }
}
}
