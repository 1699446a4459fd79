// sdfk_oracle.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Sequential / CPU restatement of the reference's hot path (praeclarum/SdfKit, pure C#), used only
// as the parity checker in tests/, by __graft_entry__.smoke() and as the `cpu_baseline` /
// `--impl reference` arm of bench.py.  Nothing under sdfkit_b200/ may call into this file.
//
// PARITY PINNING: the reference cannot run here (no .NET toolchain).  This restatement is pinned by
// the reference's own known-answer tests (tests/test_oracle_goldens.py): the nine marching-cubes
// vertex counts of Tests/SdfTests.cs:38,51 and Tests/MarchingCubesTests.cs:25-166, the sample
// values of Tests/VolumeTests.cs:92,105,134 and the depth values of Tests/RayMarcherTests.cs:21-73.
// Those goldens only reach Lewiner cases 1,2,5,8,9; the ambiguous branches (cases 3,4,6,7,10,12,13,
// face/interior tests, centre vertex), vertex positions, normals and colour renders are
// "parity unpinned" by the reference -- for them this file (a line-by-line restatement of
// MarchingCubes.cs / Cell.cs / RayMarcher.cs) is the only authority.
// A second, independent restatement written during the survey agrees with this one on every count it
// recorded (tests/golden/reference_known_answers.json "survey_probe", tests/test_golden_fixtures.py), including the triangle
// count of a white-noise field that reaches 68 decision leaves of the ambiguous cases -- the only independent pin there.
//
// Build: see oracle/Makefile (g++ -O2 -ffp-contract=off -- no FMA contraction, like RyuJIT).
//
// What is restated (reference file:line):
//   orc_sample      Voxels.SampleSdf            SdfKit/Voxels.cs:72-125  (batches, x-fastest order,
//                                               [x][y][z] z-fastest scatter, Parallel.For ~ threads)
//   orc_clip        Voxels.ClipToBounds         SdfKit/Voxels.cs:133-167
//   orc_mc_create   MarchingCubes.CreateMesh    SdfKit/MarchingCubes.cs:39-92 (+ TheBigSwitch :94-371,
//                   TestFace :376-407, TestInternal :412-546) and Cell SdfKit/Cell.cs:130-549
//   orc_mesh_transform  Mesh.Transform/Measure  SdfKit/Mesh.cs:30-64
//   orc_render / orc_render_depth   RayMarcher  SdfKit/RayMarcher.cs:45-211 (+ the VectorData.cs
//                   array passes it triggers: :164-176,428-443,490-510,715-728,735-800)
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>

#include "../sdfkit_b200/csrc/mc_luts.h"

namespace {

const signed char LUT[MCL_BLOB_SIZE] = MCL_BLOB_INIT;
const double EPS = 0.0000001;   // FLT_EPSILON of MarchingCubes.cs:37 / Cell.cs:65

struct V3 { float x, y, z; };

// ---------------------------------------------------------------------------------------------
// Cell: per-cube state + growing output lists (Cell.cs:63-143)
// ---------------------------------------------------------------------------------------------
struct Mesher {
    int nx, ny, nz;
    std::vector<V3> verts, cols, nrm;
    std::vector<int> faces;
    std::vector<int> layerA, layerB;   // faceLayer1 / faceLayer2 (Cell.cs:137-142)
    int *fl1, *fl2;
    int x = 0, y = 0, z = 0, step = 1;
    double v[8];                       // v0..v7, isovalue subtracted (Cell.cs:206-213)
    V3 c[8];
    int index = 0;
    // PrepareForAddingTriangles scratch (Cell.cs:447-499)
    double vv[8];
    V3 cc[8];
    double vg[8][3];
    // centre vertex cache (Cell.cs:501-549)
    bool v12done = false;
    double g12[3];
    float p12[3], c12[3];
    int64_t case_hist[15] = {0};
    int64_t active = 0;

    Mesher(int nx_, int ny_, int nz_) : nx(nx_), ny(ny_), nz(nz_)
    {
        layerA.assign((size_t)nx * ny * 4, -1);
        layerB.assign((size_t)nx * ny * 4, -1);
        fl1 = layerA.data();
        fl2 = layerB.data();
    }

    void new_z()   // Cell.NewZValue, Cell.cs:173-182
    {
        std::swap(fl1, fl2);
        std::fill(fl2, fl2 + (size_t)nx * ny * 4, -1);
    }

    void set_cube(double iso, int x_, int y_, int z_, int st, const float val[8], const V3 col[8])   // Cell.cs:191-233
    {
        x = x_; y = y_; z = z_; step = st;
        index = 0;
        for (int k = 0; k < 8; k++) {
            v[k] = (double)val[k] - iso;
            c[k] = col[k];
            if (v[k] > 0.0) index += 1 << k;
        }
        v12done = false;
    }

    int add_vertex(float px, float py, float pz, float r, float g, float b)   // Cell.cs:145-152
    {
        verts.push_back({px, py, pz});
        cols.push_back({r, g, b});
        nrm.push_back({0.f, 0.f, 0.f});
        return (int)verts.size() - 1;
    }

    void add_gradient(int vi, double gx, double gy, double gz)   // Cell.cs:154-155 (f32 vector add)
    {
        V3& n = nrm[vi];
        n.x = n.x + (float)gx;
        n.y = n.y + (float)gy;
        n.z = n.z + (float)gz;
    }

    void add_gradient_from(int vi, int corner, double strength)   // Cell.cs:157-158
    {
        add_gradient(vi, vg[corner][0] * strength, vg[corner][1] * strength, vg[corner][2] * strength);
    }

    // Cell.GetIndexInFacelayer (Cell.cs:371-441): slot index + which layer
    int slot_of(int e, int** layer) const
    {
        int i = nx * y + x;
        int j = 0;
        if (e < 8) {
            int h = e;
            if (e < 4) {
                *layer = fl1;
            } else {
                h = e - 4;
                *layer = fl2;
            }
            if (h == 1) { i += step; j = 1; }
            else if (h == 2) { i += nx * step; }
            else if (h == 3) { j = 1; }
        } else if (e < 12) {
            *layer = fl1;
            j = 2;
            if (e == 9) i += step;
            else if (e == 10) i += nx * step + step;
            else if (e == 11) i += nx * step;
        } else {
            *layer = fl1;
            j = 3;
        }
        return 4 * i + j;
    }

    void prepare()   // Cell.PrepareForAddingTriangles (Cell.cs:447-499); the vmax bookkeeping is dead code
    {
        static const int reorder[8] = {0, 1, 3, 2, 4, 5, 7, 6};
        for (int k = 0; k < 8; k++) {
            vv[k] = v[reorder[k]];
            cc[k] = c[reorder[k]];
        }
        const double* q = v;
        double g[8][3] = {
            {q[0] - q[1], q[0] - q[3], q[0] - q[4]},
            {q[0] - q[1], q[1] - q[2], q[1] - q[5]},
            {q[3] - q[2], q[1] - q[2], q[2] - q[6]},
            {q[3] - q[2], q[0] - q[3], q[3] - q[7]},
            {q[4] - q[5], q[4] - q[7], q[0] - q[4]},
            {q[4] - q[5], q[5] - q[6], q[1] - q[5]},
            {q[7] - q[6], q[5] - q[6], q[2] - q[6]},
            {q[7] - q[6], q[4] - q[7], q[3] - q[7]},
        };
        memcpy(vg, g, sizeof(g));
    }

    void center_vertex()   // Cell.CalculateCenterVertex (Cell.cs:501-549)
    {
        static const double ox[8] = {0, 1, 1, 0, 0, 1, 1, 0};
        static const double oy[8] = {0, 0, 1, 1, 0, 0, 1, 1};
        static const double oz[8] = {0, 0, 0, 0, 1, 1, 1, 1};
        double w[8];
        for (int k = 0; k < 8; k++) w[k] = 1.0 / (EPS + std::fabs(v[k]));
        double fx = 0.0, fy = 0.0, fz = 0.0, ff = 0.0;
        for (int k = 0; k < 8; k++) {
            fx += ox[k] * w[k];
            fy += oy[k] * w[k];
            fz += oz[k] * w[k];
            ff += w[k];
        }
        // fc = c0*(float)w0 + c1*(float)w1 + ...  (f32 vector ops, left to right)
        V3 fc = {c[0].x * (float)w[0], c[0].y * (float)w[0], c[0].z * (float)w[0]};
        for (int k = 1; k < 8; k++) {
            float wk = (float)w[k];
            fc.x = fc.x + c[k].x * wk;
            fc.y = fc.y + c[k].y * wk;
            fc.z = fc.z + c[k].z * wk;
        }
        double stp = (double)step;
        p12[0] = (float)(x + stp * fx / ff);
        p12[1] = (float)(y + stp * fy / ff);
        p12[2] = (float)(z + stp * fz / ff);
        c12[0] = (float)(fc.x / ff);
        c12[1] = (float)(fc.y / ff);
        c12[2] = (float)(fc.z / ff);
        for (int a = 0; a < 3; a++) {
            double s = w[0] * vg[0][a];
            for (int k = 1; k < 8; k++) s = s + w[k] * vg[k][a];
            g12[a] = s;
        }
        v12done = true;
    }

    void add_face_from_edge(int e)   // Cell.AddFaceFromEdgeIndex (Cell.cs:272-359)
    {
        int* layer = nullptr;
        int slot = slot_of(e, &layer);
        int vi = layer[slot];
        if (e == 12) {
            if (!v12done) center_vertex();
            if (vi < 0) {
                vi = add_vertex(p12[0], p12[1], p12[2], c12[0], c12[1], c12[2]);
                layer[slot] = vi;
            }
            faces.push_back(vi);
            add_gradient(vi, g12[0], g12[1], g12[2]);
            return;
        }
        int dx1 = LUT[MCL_edgesrelx + e * 2], dx2 = LUT[MCL_edgesrelx + e * 2 + 1];
        int dy1 = LUT[MCL_edgesrely + e * 2], dy2 = LUT[MCL_edgesrely + e * 2 + 1];
        int dz1 = LUT[MCL_edgesrelz + e * 2], dz2 = LUT[MCL_edgesrelz + e * 2 + 1];
        int i1 = dz1 * 4 + dy1 * 2 + dx1;
        int i2 = dz2 * 4 + dy2 * 2 + dx2;
        double w1 = 1.0 / (EPS + std::fabs(vv[i1]));
        double w2 = 1.0 / (EPS + std::fabs(vv[i2]));
        if (vi < 0) {
            double stp = (double)step;
            double fx = 0.0, fy = 0.0, fz = 0.0, ff = 0.0;
            fx += dx1 * w1; fy += dy1 * w1; fz += dz1 * w1; ff += w1;
            fx += dx2 * w2; fy += dy2 * w2; fz += dz2 * w2; ff += w2;
            float f1 = (float)w1, f2 = (float)w2;
            V3 col = {cc[i1].x * f1 + cc[i2].x * f2, cc[i1].y * f1 + cc[i2].y * f2, cc[i1].z * f1 + cc[i2].z * f2};
            vi = add_vertex((float)(x + stp * fx / ff), (float)(y + stp * fy / ff), (float)(z + stp * fz / ff),
                            (float)(col.x / ff), (float)(col.y / ff), (float)(col.z / ff));
            layer[slot] = vi;
        }
        faces.push_back(vi);
        add_gradient_from(vi, i1, w1);
        add_gradient_from(vi, i2, w2);
    }

    void add_tris(const signed char* row, int nt)   // Cell.AddTriangles / AddTriangles2 (Cell.cs:238-265)
    {
        prepare();
        for (int k = 0; k < 3 * nt; k++) add_face_from_edge(row[k]);
    }

    // ---- MarchingCubes.TestFace (MarchingCubes.cs:376-407)
    bool test_face(int face) const
    {
        static const int fc[7][4] = {{0, 0, 0, 0}, {0, 4, 5, 1}, {1, 5, 6, 2}, {2, 6, 7, 3}, {3, 7, 4, 0}, {0, 3, 2, 1}, {4, 7, 6, 5}};
        int af = face < 0 ? -face : face;
        double A = 0.0, B = 0.0, C = 0.0, D = 0.0;
        if (af >= 1 && af <= 6) {
            A = v[fc[af][0]]; B = v[fc[af][1]]; C = v[fc[af][2]]; D = v[fc[af][3]];
        }
        double acbd = A * C - B * D;
        if (acbd > -EPS && acbd < EPS) return face >= 0;
        return face * A * acbd >= 0;
    }

    // ---- MarchingCubes.TestInternal (MarchingCubes.cs:412-546)
    bool test_internal(int cas, int config, int subconfig, int s) const
    {
        double t, At = 0.0, Bt = 0.0, Ct = 0.0, Dt = 0.0;
        if (cas == 4 || cas == 10) {
            double a = (v[4] - v[0]) * (v[6] - v[2]) - (v[7] - v[3]) * (v[5] - v[1]);
            double b = v[2] * (v[4] - v[0]) + v[0] * (v[6] - v[2]) - v[1] * (v[7] - v[3]) - v[3] * (v[5] - v[1]);
            t = -b / (2 * a + EPS);
            if (t < 0 || t > 1) return s > 0;
            At = v[0] + (v[4] - v[0]) * t;
            Bt = v[3] + (v[7] - v[3]) * t;
            Ct = v[2] + (v[6] - v[2]) * t;
            Dt = v[1] + (v[5] - v[1]) * t;
        } else if (cas == 6 || cas == 7 || cas == 12 || cas == 13) {
            int edge = -1;
            if (cas == 6) edge = LUT[MCL_test6 + config * 3 + 2];
            else if (cas == 7) edge = LUT[MCL_test7 + config * 5 + 4];
            else if (cas == 12) edge = LUT[MCL_test12 + config * 4 + 3];
            else edge = LUT[MCL_tiling13_5_1 + (config * 4 + subconfig) * 18 + 0];
            // {t numerator corner, t other corner, B from,to, C from,to, D from,to}  (MarchingCubes.cs:440-511)
            static const int tab[12][8] = {
                {0, 1, 3, 2, 7, 6, 4, 5}, {1, 2, 0, 3, 4, 7, 5, 6}, {2, 3, 1, 0, 5, 4, 6, 7}, {3, 0, 2, 1, 6, 5, 7, 4},
                {4, 5, 7, 6, 3, 2, 0, 1}, {5, 6, 4, 7, 0, 3, 1, 2}, {6, 7, 5, 4, 1, 0, 2, 3}, {7, 4, 6, 5, 2, 1, 3, 0},
                {0, 4, 3, 7, 2, 6, 1, 5}, {1, 5, 0, 4, 3, 7, 2, 6}, {2, 6, 1, 5, 0, 4, 3, 7}, {3, 7, 2, 6, 1, 5, 0, 4}};
            if (edge >= 0 && edge < 12) {
                const int* r = tab[edge];
                t = v[r[0]] / (v[r[0]] - v[r[1]] + EPS);
                At = 0;
                Bt = v[r[2]] + (v[r[3]] - v[r[2]]) * t;
                Ct = v[r[4]] + (v[r[5]] - v[r[4]]) * t;
                Dt = v[r[6]] + (v[r[7]] - v[r[6]]) * t;
            }
        }
        int test = 0;
        if (At >= 0) test += 1;
        if (Bt >= 0) test += 2;
        if (Ct >= 0) test += 4;
        if (Dt >= 0) test += 8;
        switch (test) {
        case 5: if (At * Ct - Bt * Dt < EPS) return s > 0; break;
        case 10: if (At * Ct - Bt * Dt >= EPS) return s > 0; break;
        case 7: case 11: case 13: case 14: case 15: return s < 0;
        default: return s > 0;   // 0,1,2,3,4,6,8,9,12
        }
        return s < 0;
    }

#define ROW2(name, cfg) (&LUT[MCL_##name + (cfg) * MCL_##name##_D1])
#define ROW3(name, cfg, sub) (&LUT[MCL_##name + ((cfg) * MCL_##name##_D1 + (sub)) * MCL_##name##_D2])

    // MarchingCubes.TheBigSwitch (MarchingCubes.cs:94-371)
    void big_switch(int cas, int cfg)
    {
        int sub = 0;
        switch (cas) {
        case 1: add_tris(ROW2(tiling1, cfg), 1); break;
        case 2: add_tris(ROW2(tiling2, cfg), 2); break;
        case 3:
            if (test_face(LUT[MCL_test3 + cfg])) add_tris(ROW2(tiling3_2, cfg), 4);
            else add_tris(ROW2(tiling3_1, cfg), 2);
            break;
        case 4:
            if (test_internal(cas, cfg, sub, LUT[MCL_test4 + cfg])) add_tris(ROW2(tiling4_1, cfg), 2);
            else add_tris(ROW2(tiling4_2, cfg), 6);
            break;
        case 5: add_tris(ROW2(tiling5, cfg), 3); break;
        case 6:
            if (test_face(LUT[MCL_test6 + cfg * 3 + 0])) add_tris(ROW2(tiling6_2, cfg), 5);
            else if (test_internal(cas, cfg, sub, LUT[MCL_test6 + cfg * 3 + 1])) add_tris(ROW2(tiling6_1_1, cfg), 3);
            else add_tris(ROW2(tiling6_1_2, cfg), 9);
            break;
        case 7:
            if (test_face(LUT[MCL_test7 + cfg * 5 + 0])) sub += 1;
            if (test_face(LUT[MCL_test7 + cfg * 5 + 1])) sub += 2;
            if (test_face(LUT[MCL_test7 + cfg * 5 + 2])) sub += 4;
            switch (sub) {
            case 0: add_tris(ROW2(tiling7_1, cfg), 3); break;
            case 1: add_tris(ROW3(tiling7_2, cfg, 0), 5); break;
            case 2: add_tris(ROW3(tiling7_2, cfg, 1), 5); break;
            case 3: add_tris(ROW3(tiling7_3, cfg, 0), 9); break;
            case 4: add_tris(ROW3(tiling7_2, cfg, 2), 5); break;
            case 5: add_tris(ROW3(tiling7_3, cfg, 1), 9); break;
            case 6: add_tris(ROW3(tiling7_3, cfg, 2), 9); break;
            case 7:
                if (test_internal(cas, cfg, sub, LUT[MCL_test7 + cfg * 5 + 3])) add_tris(ROW2(tiling7_4_2, cfg), 9);
                else add_tris(ROW2(tiling7_4_1, cfg), 5);
                break;
            }
            break;
        case 8: add_tris(ROW2(tiling8, cfg), 2); break;
        case 9: add_tris(ROW2(tiling9, cfg), 4); break;
        case 10:
            if (test_face(LUT[MCL_test10 + cfg * 3 + 0])) {
                if (test_face(LUT[MCL_test10 + cfg * 3 + 1])) add_tris(ROW2(tiling10_1_1_, cfg), 4);
                else add_tris(ROW2(tiling10_2, cfg), 8);
            } else {
                if (test_face(LUT[MCL_test10 + cfg * 3 + 1])) add_tris(ROW2(tiling10_2_, cfg), 8);
                else if (test_internal(cas, cfg, sub, LUT[MCL_test10 + cfg * 3 + 2])) add_tris(ROW2(tiling10_1_1, cfg), 4);
                else add_tris(ROW2(tiling10_1_2, cfg), 8);
            }
            break;
        case 11: add_tris(ROW2(tiling11, cfg), 4); break;
        case 12:
            if (test_face(LUT[MCL_test12 + cfg * 4 + 0])) {
                if (test_face(LUT[MCL_test12 + cfg * 4 + 1])) add_tris(ROW2(tiling12_1_1_, cfg), 4);
                else add_tris(ROW2(tiling12_2, cfg), 8);
            } else {
                if (test_face(LUT[MCL_test12 + cfg * 4 + 1])) add_tris(ROW2(tiling12_2_, cfg), 8);
                else if (test_internal(cas, cfg, sub, LUT[MCL_test12 + cfg * 4 + 2])) add_tris(ROW2(tiling12_1_1, cfg), 4);
                else add_tris(ROW2(tiling12_1_2, cfg), 8);
            }
            break;
        case 13: {
            for (int k = 0; k < 6; k++)
                if (test_face(LUT[MCL_test13 + cfg * 7 + k])) sub += 1 << k;
            sub = LUT[MCL_subconfig13 + sub];
            if (sub < 0) break;   // "Impossible case 13?" (MarchingCubes.cs:364-366): no triangles
            if (sub == 0) add_tris(ROW2(tiling13_1, cfg), 4);
            else if (sub <= 6) add_tris(ROW3(tiling13_2, cfg, sub - 1), 6);
            else if (sub <= 18) add_tris(ROW3(tiling13_3, cfg, sub - 7), 10);
            else if (sub <= 22) add_tris(ROW3(tiling13_4, cfg, sub - 19), 12);
            else if (sub <= 26) {
                int s5 = sub - 23;
                if (test_internal(cas, cfg, s5, LUT[MCL_test13 + cfg * 7 + 6])) add_tris(ROW3(tiling13_5_1, cfg, s5), 6);
                else add_tris(ROW3(tiling13_5_2, cfg, s5), 10);
            } else if (sub <= 38) add_tris(ROW3(tiling13_3_, cfg, sub - 27), 10);
            else if (sub <= 44) add_tris(ROW3(tiling13_2_, cfg, sub - 39), 6);
            else if (sub == 45) add_tris(ROW2(tiling13_1_, cfg), 4);
            break;   // anything else: "Impossible case 13?" -- no triangles
        }
        case 14: add_tris(ROW2(tiling14, cfg), 4); break;
        }
    }
};

struct OrcMesh {
    std::vector<V3> verts, cols, nrm;
    std::vector<int> tris;
    float aabb[6] = {0, 0, 0, 0, 0, 0};
    int64_t case_hist[15];
    int64_t active;

    void measure()   // Mesh.Measure (Mesh.cs:30-45): Vector3.Min/Max = (a<b)?a:b / (a>b)?a:b
    {
        if (verts.empty()) return;
        V3 mn = verts[0], mx = verts[0];
        for (size_t i = 1; i < verts.size(); i++) {
            const V3& q = verts[i];
            mn.x = (mn.x < q.x) ? mn.x : q.x; mn.y = (mn.y < q.y) ? mn.y : q.y; mn.z = (mn.z < q.z) ? mn.z : q.z;
            mx.x = (mx.x > q.x) ? mx.x : q.x; mx.y = (mx.y > q.y) ? mx.y : q.y; mx.z = (mx.z > q.z) ? mx.z : q.z;
        }
        aabb[0] = mn.x; aabb[1] = mn.y; aabb[2] = mn.z; aabb[3] = mx.x; aabb[4] = mx.y; aabb[5] = mx.z;
    }
};

inline V3 normalize_div(V3 a)   // Vector3.Normalize = value / value.Length()
{
    float len = sqrtf((a.x * a.x + a.y * a.y) + a.z * a.z);
    return {a.x / len, a.y / len, a.z / len};
}

}  // namespace

extern "C" {

typedef void (*orc_sdf_fn)(const float* pts, float* out, int n);   // the Sdf delegate, Sdf.cs:8
typedef void (*orc_progress_fn)(float, void*);

// Voxels.SampleSdf (Voxels.cs:72-125).  values: float[nx][ny][nz]; colors: float[nx][ny][nz][3] (C# layout).
// batch_sizes (optional, may be NULL) receives the length of every batch, for Tests/VolumeTests.cs:116-118.
void orc_sample(orc_sdf_fn sdf, const float mn[3], const float mx[3], int nx, int ny, int nz, int batch, int threads,
                float* values, float* colors, int* batch_sizes)
{
    const float DX = nx >= 1 ? (mx[0] - mn[0]) / nx : 0.0f;   // Voxels.cs:32-34
    const float DY = ny >= 1 ? (mx[1] - mn[1]) / ny : 0.0f;
    const float DZ = nz >= 1 ? (mx[2] - mn[2]) / nz : 0.0f;
    const float m0 = mn[0] + 0.5f * DX, m1 = mn[1] + 0.5f * DY, m2 = mn[2] + 0.5f * DZ;   // Voxels.cs:81
    const int ntotal = nx * ny * nz;
    const int nbatch = (ntotal + batch - 1) / batch;
    std::atomic<int> next(0);
    auto worker = [&]() {
        std::vector<float> pos((size_t)batch * 3), val((size_t)batch * 4);   // thread-local scratch, Voxels.cs:88-93
        for (;;) {
            int ib = next.fetch_add(1);
            if (ib >= nbatch) break;
            int s = ib * batch, e = std::min(ntotal, s + batch);
            for (int i = s; i < e; i++) {
                int ix = i % nx, iy = (i / nx) % ny, iz = i / (nx * ny);
                pos[(size_t)(i - s) * 3 + 0] = m0 + ix * DX;
                pos[(size_t)(i - s) * 3 + 1] = m1 + iy * DY;
                pos[(size_t)(i - s) * 3 + 2] = m2 + iz * DZ;
            }
            sdf(pos.data(), val.data(), e - s);
            if (batch_sizes) batch_sizes[ib] = e - s;
            for (int i = s; i < e; i++) {
                int ix = i % nx, iy = (i / nx) % ny, iz = i / (nx * ny);
                size_t o = ((size_t)ix * ny + iy) * nz + iz;
                const float* q = &val[(size_t)(i - s) * 4];
                values[o] = q[3];
                colors[o * 3 + 0] = q[0];
                colors[o * 3 + 1] = q[1];
                colors[o * 3 + 2] = q[2];
            }
        }
    };
    if (threads <= 1) {
        worker();
    } else {
        std::vector<std::thread> pool;
        for (int t = 0; t < threads; t++) pool.emplace_back(worker);
        for (auto& t : pool) t.join();
    }
}

// Voxels.ClipToBounds (Voxels.cs:133-167): six faces := Size.X / NX
void orc_clip(float* values, const float mn[3], const float mx[3], int nx, int ny, int nz)
{
    const float outside = (mx[0] - mn[0]) / nx;
    auto at = [&](int ix, int iy, int iz) -> float& { return values[((size_t)ix * ny + iy) * nz + iz]; };
    for (int iy = 0; iy < ny; iy++)
        for (int iz = 0; iz < nz; iz++) { at(0, iy, iz) = outside; at(nx - 1, iy, iz) = outside; }
    for (int ix = 0; ix < nx; ix++)
        for (int iz = 0; iz < nz; iz++) { at(ix, 0, iz) = outside; at(ix, ny - 1, iz) = outside; }
    for (int ix = 0; ix < nx; ix++)
        for (int iy = 0; iy < ny; iy++) { at(ix, iy, 0) = outside; at(ix, iy, nz - 1) = outside; }
}

// MarchingCubes.CreateMesh up to (not including) the final transform (MarchingCubes.cs:39-84).
// cell_index / cell_ntris (optional): per visited cell, in visit order (z outer, y, x inner).
void* orc_mc_create(const float* values, const float* colors, int nx, int ny, int nz, float iso, int step,
                    orc_progress_fn progress, void* user, unsigned char* cell_index, unsigned char* cell_ntris)
{
    Mesher cell(nx, ny, nz);
    const int nxb = nx - 2 * step, nyb = ny - 2 * step, nzb = nz - 2 * step;
    auto off = [&](int ix, int iy, int iz) { return ((size_t)ix * ny + iy) * nz + iz; };
    size_t visit = 0;
    int z = -step;
    while (z < nzb) {
        z += step;
        int zs = z + step;
        cell.new_z();
        int y = -step;
        while (y < nyb) {
            y += step;
            int ys = y + step;
            int x = -step;
            while (x < nxb) {
                x += step;
                int xs = x + step;
                const size_t o[8] = {off(x, y, z), off(xs, y, z), off(xs, ys, z), off(x, ys, z),
                                     off(x, y, zs), off(xs, y, zs), off(xs, ys, zs), off(x, ys, zs)};
                float val[8];
                V3 col[8];
                for (int k = 0; k < 8; k++) {
                    val[k] = values[o[k]];
                    col[k] = {colors[o[k] * 3], colors[o[k] * 3 + 1], colors[o[k] * 3 + 2]};
                }
                cell.set_cube((double)iso, x, y, z, step, val, col);
                int cas = LUT[MCL_cases + cell.index * 2];
                size_t before = cell.faces.size();
                if (cas > 0) {
                    cell.case_hist[cas]++;
                    cell.active++;
                    cell.big_switch(cas, LUT[MCL_cases + cell.index * 2 + 1]);
                }
                if (cell_index) cell_index[visit] = (unsigned char)cell.index;
                if (cell_ntris) cell_ntris[visit] = (unsigned char)((cell.faces.size() - before) / 3);
                visit++;
            }
        }
        if (progress) progress((float)z / nzb, user);
    }
    OrcMesh* m = new OrcMesh();
    m->verts = std::move(cell.verts);
    m->cols = std::move(cell.cols);
    m->tris = std::move(cell.faces);
    m->nrm.resize(cell.nrm.size());
    for (size_t i = 0; i < cell.nrm.size(); i++) {   // Cell.NegativeNormals (Cell.cs:97-109)
        V3 n = normalize_div(cell.nrm[i]);
        m->nrm[i] = {-n.x, -n.y, -n.z};
    }
    memcpy(m->case_hist, cell.case_hist, sizeof(cell.case_hist));
    m->active = cell.active;
    m->measure();
    return m;
}

// Mesh.Transform (Mesh.cs:47-64).  M: the 4x4 transform, N: the "normal transform" (transpose of the inverse
// of M with its translation cleared) -- both row-major System.Numerics matrices, computed by the caller.
void orc_mesh_transform(void* h, const float M[16], const float N[16])
{
    OrcMesh* m = (OrcMesh*)h;
    for (size_t i = 0; i < m->verts.size(); i++) {
        V3 p = m->verts[i], n = m->nrm[i];
        V3 tp = {p.x * M[0] + p.y * M[4] + p.z * M[8] + M[12], p.x * M[1] + p.y * M[5] + p.z * M[9] + M[13],
                 p.x * M[2] + p.y * M[6] + p.z * M[10] + M[14]};
        V3 tn = {n.x * N[0] + n.y * N[4] + n.z * N[8], n.x * N[1] + n.y * N[5] + n.z * N[9],
                 n.x * N[2] + n.y * N[6] + n.z * N[10]};
        m->verts[i] = tp;
        m->nrm[i] = normalize_div(tn);
    }
    m->measure();
}

void orc_mesh_counts(void* h, int64_t* nverts, int64_t* ntris, int64_t* active, int64_t case_hist[15])
{
    OrcMesh* m = (OrcMesh*)h;
    if (nverts) *nverts = (int64_t)m->verts.size();
    if (ntris) *ntris = (int64_t)m->tris.size() / 3;
    if (active) *active = m->active;
    if (case_hist) memcpy(case_hist, m->case_hist, sizeof(m->case_hist));
}

void orc_mesh_export(void* h, float* verts, float* colors, float* normals, int* tris, float aabb[6])
{
    OrcMesh* m = (OrcMesh*)h;
    if (verts && !m->verts.empty()) memcpy(verts, m->verts.data(), m->verts.size() * sizeof(V3));
    if (colors && !m->cols.empty()) memcpy(colors, m->cols.data(), m->cols.size() * sizeof(V3));
    if (normals && !m->nrm.empty()) memcpy(normals, m->nrm.data(), m->nrm.size() * sizeof(V3));
    if (tris && !m->tris.empty()) memcpy(tris, m->tris.data(), m->tris.size() * sizeof(int));
    if (aabb) memcpy(aabb, m->aabb, sizeof(m->aabb));
}

void orc_mesh_free(void* h) { delete (OrcMesh*)h; }

// ---------------------------------------------------------------------------------------------
// Ray marcher (RayMarcher.cs).  The camera matrices are computed by the caller (host code with the
// System.Numerics restatement): cam_pos = translation of inverse(view), ivp = inverse(view*proj).
// ---------------------------------------------------------------------------------------------

// RayMarcher.GetCameraRays per-pixel part (RayMarcher.cs:111-125)
static void camera_rays(int w, int h, const float cam[3], const float ivp[16], float* rd)
{
    size_t k = 0;
    for (int j = 0; j < h; j++) {
        float y = 1.0f - 2.0f * (float)j / (h - 1);
        for (int i = 0; i < w; i++) {
            float x = -1.0f + 2.0f * (float)i / (w - 1);
            // Vector4.Transform((x,y,0,1), ivp)
            float tx = x * ivp[0] + y * ivp[4] + 0.0f * ivp[8] + 1.0f * ivp[12];
            float ty = x * ivp[1] + y * ivp[5] + 0.0f * ivp[9] + 1.0f * ivp[13];
            float tz = x * ivp[2] + y * ivp[6] + 0.0f * ivp[10] + 1.0f * ivp[14];
            float tw = x * ivp[3] + y * ivp[7] + 0.0f * ivp[11] + 1.0f * ivp[15];
            V3 d = {tx / tw - cam[0], ty / tw - cam[1], tz / tw - cam[2]};
            d = normalize_div(d);
            rd[k++] = d.x; rd[k++] = d.y; rd[k++] = d.z;
        }
    }
}

// RayMarcher.Scene (RayMarcher.cs:206-211): SdfEx.Sample with maxDegreeOfParallelism 1 (Sdf.cs:26-33)
static void scene(orc_sdf_fn sdf, const float* pts, float* out, int n, int batch)
{
    for (int s = 0; s < n; s += batch) sdf(pts + (size_t)s * 3, out + (size_t)s * 4, std::min(batch, n - s));
}

// RayMarcher.Render(frag, ro, rd) on one row band (RayMarcher.cs:131-204), whole-band array passes like the reference
static void render_band(orc_sdf_fn sdf, int npix, const float cam[3], const float* rd, float nearp, float farp,
                        int iters, int batch, float* frag)
{
    std::vector<float> depth(npix, nearp - 0.1f), diffuse((size_t)npix * 3, 0.0f);
    std::vector<float> pos((size_t)npix * 3), sd((size_t)npix * 4);
    for (int it = 0; it < iters; it++) {
        for (int p = 0; p < npix; p++)   // MulAdd(rayDir, depth, rayOrigin)  VectorData.cs:789-797
            for (int a = 0; a < 3; a++) pos[(size_t)p * 3 + a] = rd[(size_t)p * 3 + a] * depth[p] + cam[a];
        scene(sdf, pos.data(), sd.data(), npix, batch);
        for (int p = 0; p < npix; p++) depth[p] += sd[(size_t)p * 4 + 3];
        if (it == iters - 1)
            for (int p = 0; p < npix; p++)
                for (int a = 0; a < 3; a++) diffuse[(size_t)p * 3 + a] += sd[(size_t)p * 4 + a];
    }
    std::vector<float> surf((size_t)npix * 3);
    for (int p = 0; p < npix; p++)   // rayOrigin + rayDir*depth
        for (int a = 0; a < 3; a++) surf[(size_t)p * 3 + a] = cam[a] + rd[(size_t)p * 3 + a] * depth[p];
    // DistanceGradient (RayMarcher.cs:164-204): taps +x,+y,+z,-x,-y,-z at 1e-5
    const float go = 1e-5f;
    const float offs[6][3] = {{go * 1.0f, go * 0.0f, go * 0.0f}, {go * 0.0f, go * 1.0f, go * 0.0f}, {go * 0.0f, go * 0.0f, go * 1.0f},
                              {-go * 1.0f, -go * 0.0f, -go * 0.0f}, {-go * 0.0f, -go * 1.0f, -go * 0.0f}, {-go * 0.0f, -go * 0.0f, -go * 1.0f}};
    std::vector<float> tap[6];
    for (int k = 0; k < 6; k++) {
        for (int p = 0; p < npix; p++)
            for (int a = 0; a < 3; a++) pos[(size_t)p * 3 + a] = surf[(size_t)p * 3 + a] + offs[k][a];
        scene(sdf, pos.data(), sd.data(), npix, batch);
        tap[k].resize(npix);
        for (int p = 0; p < npix; p++) tap[k][p] = sd[(size_t)p * 4 + 3];
    }
    for (int p = 0; p < npix; p++) {
        float nxv = tap[0][p] - tap[3][p], nyv = tap[1][p] - tap[4][p], nzv = tap[2][p] - tap[5][p];
        float len = sqrtf(nxv * nxv + nyv * nyv + nzv * nzv);   // NormalizeInplace (VectorData.cs:490-510)
        if (len > 0) { float r = 1.0f / len; nxv = nxv * r; nyv = nyv * r; nzv = nzv * r; }
        float lx = 5.0f - surf[(size_t)p * 3], ly = 5.0f - surf[(size_t)p * 3 + 1], lz = 10.0f - surf[(size_t)p * 3 + 2];
        float ll = sqrtf(lx * lx + ly * ly + lz * lz);
        if (ll > 0) { float r = 1.0f / ll; lx = lx * r; ly = ly * r; lz = lz * r; }
        float dv = nxv * lx + nyv * ly + nzv * lz;      // DotInplace (VectorData.cs:464-475)
        dv = (dv != dv) ? dv : ((dv > 0.0f) ? dv : 0.0f);   // MaxInplace: MathF.Max(v, 0) (NaN propagates)
        float mask = depth[p] > farp ? 1.0f : 0.0f;     // FloatData operator > (VectorData.cs:181-191)
        float notmask = mask == 0.0f ? 1.0f : 0.0f;     // NotInplace
        const float bgc[3] = {0.5f, 0.75f, 1.0f};
        for (int a = 0; a < 3; a++) {
            float lighting = dv * diffuse[(size_t)p * 3 + a] + 0.1f;   // MulAdd(FloatData, Vec3Data, float)
            float bg = mask * bgc[a];
            float fg = lighting * notmask + bg;                        // MulAdd(Vec3Data, FloatData, Vec3Data)
            frag[(size_t)p * 3 + a] += fg;
        }
    }
}

// RayMarcher.Render (RayMarcher.cs:45-64): rays single-threaded, then `bands` row bands in parallel.
void orc_render(orc_sdf_fn sdf, int w, int h, const float cam[3], const float ivp[16], float nearp, float farp, int iters,
                int batch, int bands, float* rgb)
{
    std::vector<float> rd((size_t)w * h * 3);
    camera_rays(w, h, cam, ivp, rd.data());
    memset(rgb, 0, (size_t)w * h * 3 * sizeof(float));
    if (bands < 1) bands = 1;
    int bandh = (h + bands - 1) / bands;   // PartitionVertically (VectorData.cs:512-526)
    std::vector<std::thread> pool;
    int y = 0;
    for (int b = 0; b < bands; b++) {
        int hh = std::min(bandh, h - y);
        if (hh <= 0) break;
        const float* rdp = rd.data() + (size_t)y * w * 3;
        float* fp = rgb + (size_t)y * w * 3;
        pool.emplace_back([=]() { render_band(sdf, w * hh, cam, rdp, nearp, farp, iters, batch, fp); });
        y += hh;
    }
    for (auto& t : pool) t.join();
}

// RayMarcher.RenderDepth (RayMarcher.cs:69-93): single-threaded
void orc_render_depth(orc_sdf_fn sdf, int w, int h, const float cam[3], const float ivp[16], float nearp, int iters,
                      int batch, float* depth)
{
    const int npix = w * h;
    std::vector<float> rd((size_t)npix * 3), pos((size_t)npix * 3), sd((size_t)npix * 4);
    camera_rays(w, h, cam, ivp, rd.data());
    for (int p = 0; p < npix; p++) depth[p] = nearp - 0.1f;
    for (int it = 0; it < iters; it++) {
        for (int p = 0; p < npix; p++)   // rayDir*depth, then += rayOrigin
            for (int a = 0; a < 3; a++) pos[(size_t)p * 3 + a] = rd[(size_t)p * 3 + a] * depth[p] + cam[a];
        scene(sdf, pos.data(), sd.data(), npix, batch);
        for (int p = 0; p < npix; p++) depth[p] += sd[(size_t)p * 4 + 3];
    }
}

}  // extern "C"
