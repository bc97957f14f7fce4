// Exercises the C++ host layer (include/brush_b200.hpp).  Modes:
//   uniforms            stdin: one camera per line -> stdout: the BgCamera fields as hex floats   (no GPU needed)
//   fov                 stdin: "focal pixels model p0..p7" -> fov and the focal it maps back to    (no GPU needed)
//   errors              exceptions for invalid arguments                                           (no GPU needed)
//   render IN OUT       forward + backward of a scene file written by tests/test_cpp_host.py       (GPU)
//   train IN STEPS      SplatTrainer::step on the same kind of file (+ packed ground truth)         (GPU)
//   bounds              BoundingBox::median_size on the cases of bounding_box.rs:36-58               (no GPU needed)
//   refine IN STEPS     STEPS x SplatTrainer::step_views (one view, one device), refine, one more step (GPU)
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#include "brush_b200.hpp"

using namespace brush_b200;

static Camera read_camera(std::istream &in, uint32_t &w, uint32_t &h) {
    Camera c;
    uint32_t model;
    in >> c.position[0] >> c.position[1] >> c.position[2] >> c.rotation[0] >> c.rotation[1] >> c.rotation[2] >> c.rotation[3] >>
        c.fov_x >> c.fov_y >> c.center_uv[0] >> c.center_uv[1] >> model;
    c.model = (CameraModel)model;
    for (int i = 0; i < 8; i++) in >> c.model_params[i];
    in >> w >> h;
    return c;
}

static void print_uniforms(const BgCamera &u) {
    for (int i = 0; i < 12; i++) std::printf("%a ", u.viewmat[i]);
    std::printf("%a %a %a %a %a %a %a %a %a %a %a %a %u", u.fx, u.fy, u.cx, u.cy, u.cam_pos[0], u.cam_pos[1], u.cam_pos[2],
                u.lim_pos_x, u.lim_pos_y, u.lim_neg_x, u.lim_neg_y, u.half_max_render_fov, u.camera_model);
    for (int i = 0; i < 8; i++) std::printf(" %a", u.model_params[i]);
    std::printf("\n");
}

template <typename T>
static std::vector<T> read_vec(std::ifstream &f, size_t n) {
    std::vector<T> v(n);
    f.read(reinterpret_cast<char *>(v.data()), n * sizeof(T));
    return v;
}
template <typename T>
static void write_vec(std::ofstream &f, const std::vector<T> &v) { f.write(reinterpret_cast<const char *>(v.data()), v.size() * sizeof(T)); }

int main(int argc, char **argv) {
    const std::string mode = argc > 1 ? argv[1] : "";
    try {
        if (mode == "uniforms") {
            std::string line;
            while (std::getline(std::cin, line)) {
                if (line.empty()) continue;
                std::istringstream ss(line);
                uint32_t w, h;
                Camera c = read_camera(ss, w, h);
                print_uniforms(make_uniforms(c, w, h));
            }
            return 0;
        }
        if (mode == "fov") {
            double focal; uint32_t px, model; float p[8];
            while (std::cin >> focal >> px >> model) {
                for (float &v : p) std::cin >> v;
                double fov = focal_to_fov(focal, px, (CameraModel)model, p);
                std::printf("%.17g %.17g\n", fov, fov_to_focal(fov, px, (CameraModel)model, p));
            }
            return 0;
        }
        if (mode == "errors") {
            int caught = 0;
            try { make_uniforms(Camera{}, 0, 16); } catch (const Error &e) { caught += e.status == BG_ERR_INVALID; }
            try { Context c(0, 0, 64, 64); } catch (const Error &e) { caught += e.status == BG_ERR_INVALID; }
            try { check(bg_ctx_destroy(nullptr), "destroy"); } catch (const Error &e) { caught += e.status == BG_ERR_NULL; }
            std::printf("caught %d\n", caught);
            // the entry points behind step_views / refine / DpComm reject null handles before touching a device
            int more = 0;
            try { check(bg_train_step_views(nullptr, nullptr, nullptr, nullptr), "step_views"); } catch (const Error &e) { more += e.status == BG_ERR_NULL; }
            try { BgRefineStats rs; check(bg_refine(nullptr, nullptr, nullptr, &rs), "refine"); } catch (const Error &e) { more += e.status == BG_ERR_NULL; }
            try { BgDpComm *h = nullptr; uint8_t id[128] = {0}; check(bg_dp_comm_create(nullptr, id, 0, 1, &h), "dp"); } catch (const Error &e) { more += e.status == BG_ERR_NULL; }
            std::printf("caught_more %d\n", more);
            return caught == 3 && more == 3 ? 0 : 1;
        }
        if (mode == "bounds") {   // bounding_box.rs:36-58
            BoundingBox a, b, c;
            a.extent[0] = NAN; a.extent[1] = 2.0f; a.extent[2] = 3.0f;
            b.extent[0] = b.extent[1] = b.extent[2] = NAN;
            c.extent[0] = 1.0f; c.extent[1] = 2.0f; c.extent[2] = 3.0f;   // from_min_max(-1, (1, 3, 5))
            std::printf("%.9g %.9g %.9g\n", a.median_size(), b.median_size(), c.median_size());
            return 0;
        }
        if (mode == "render" && argc == 4) {
            std::ifstream f(argv[2], std::ios::binary);
            uint32_t hdr[6];   // n k w h mip pass
            f.read(reinterpret_cast<char *>(hdr), sizeof(hdr));
            const uint32_t n = hdr[0], k = hdr[1], w = hdr[2], h = hdr[3];
            std::string camline;
            { uint32_t len; f.read(reinterpret_cast<char *>(&len), 4); camline.resize(len); f.read(&camline[0], len); }
            std::istringstream ss(camline);
            uint32_t cw, ch;
            Camera cam = read_camera(ss, cw, ch);
            float bg[3];
            f.read(reinterpret_cast<char *>(bg), sizeof(bg));
            auto tr = read_vec<float>(f, (size_t)n * 10), sh = read_vec<float>(f, (size_t)n * k * 3), op = read_vec<float>(f, n);
            auto v_out = read_vec<float>(f, (size_t)w * h * 4);
            Context ctx(0, n, w, h);
            DeviceBuffer<float> d_tr(tr.size()), d_sh(sh.size()), d_op(op.size()), d_vo(v_out.size());
            d_tr.upload(tr.data(), tr.size()); d_sh.upload(sh.data(), sh.size()); d_op.upload(op.data(), op.size());
            d_vo.upload(v_out.data(), v_out.size());
            RenderOutput out = render(ctx, nullptr, cam, w, h, d_tr.data(), d_sh.data(), d_op.data(), n, k,
                                      hdr[4] ? SplatRenderMode::Mip : SplatRenderMode::Default, bg, (RasterPass)hdr[5]);
            DeviceBuffer<float> vc = rasterize_bwd(ctx, nullptr, out, d_vo.data(), bg, hdr[5] == 2);
            SplatGrads g = project_bwd(ctx, nullptr, out, d_tr.data(), d_sh.data(), d_op.data(), vc.data());
            std::vector<float> img((size_t)w * h * 4), vt((size_t)n * 10), vsh((size_t)n * k * 3), vo(n), vis(n);
            out.out_img_f32.download(img.data(), img.size());
            g.v_transforms.download(vt.data(), vt.size());
            g.v_coeffs.download(vsh.data(), vsh.size());
            g.v_raw_opac.download(vo.data(), vo.size());
            out.visible.download(vis.data(), vis.size());
            std::ofstream o(argv[3], std::ios::binary);
            uint32_t counts[2] = {out.num_visible(), out.num_intersections()};
            o.write(reinterpret_cast<const char *>(counts), sizeof(counts));
            write_vec(o, img); write_vec(o, vt); write_vec(o, vsh); write_vec(o, vo); write_vec(o, vis);
            std::printf("V %u I %u arena %llu\n", counts[0], counts[1], (unsigned long long)ctx.arena_bytes());
            return 0;
        }
        if (mode == "train" && argc == 4) {   // scene file as for `render`, then the packed ground truth [h,w] u32; argv[3]: steps
            std::ifstream f(argv[2], std::ios::binary);
            uint32_t hdr[6];
            f.read(reinterpret_cast<char *>(hdr), sizeof(hdr));
            const uint32_t n = hdr[0], k = hdr[1], w = hdr[2], h = hdr[3];
            std::string camline;
            { uint32_t len; f.read(reinterpret_cast<char *>(&len), 4); camline.resize(len); f.read(&camline[0], len); }
            std::istringstream ss(camline);
            uint32_t cw, ch;
            Camera cam = read_camera(ss, cw, ch);
            float bg[3];
            f.read(reinterpret_cast<char *>(bg), sizeof(bg));
            auto tr = read_vec<float>(f, (size_t)n * 10), sh = read_vec<float>(f, (size_t)n * k * 3), op = read_vec<float>(f, n);
            auto skip = read_vec<float>(f, (size_t)w * h * 4);
            auto gt = read_vec<uint32_t>(f, (size_t)w * h);
            float median_scale;
            f.read(reinterpret_cast<char *>(&median_scale), 4);
            Context ctx(0, n, w, h);
            DeviceBuffer<float> d_tr(tr.size()), d_sh(sh.size()), d_op(op.size());
            DeviceBuffer<uint32_t> d_gt(gt.size());
            d_tr.upload(tr.data(), tr.size()); d_sh.upload(sh.data(), sh.size()); d_op.upload(op.data(), op.size());
            d_gt.upload(gt.data(), gt.size());
            TrainConfig cfg;
            cfg.total_train_iters = 1000;
            cfg.seed = 7;
            SplatTrainer trainer(cfg, n, k, median_scale);
            const int steps = std::atoi(argv[3]);
            for (int i = 0; i < steps; i++) {
                const float *loss_dev = trainer.step(ctx, nullptr, cam, d_gt.data(), w, h, d_tr.data(), d_sh.data(), d_op.data());
                float loss;
                check_cuda(cudaMemcpy(&loss, loss_dev, 4, cudaMemcpyDeviceToHost), "loss readback");
                std::printf("loss %.9g\n", loss);
            }
            return 0;
        }
        if (mode == "refine" && argc == 4) {   // scene file as for `train`; argv[3]: steps before the refine
            std::ifstream f(argv[2], std::ios::binary);
            uint32_t hdr[6];
            f.read(reinterpret_cast<char *>(hdr), sizeof(hdr));
            const uint32_t n = hdr[0], k = hdr[1], w = hdr[2], h = hdr[3];
            std::string camline;
            { uint32_t len; f.read(reinterpret_cast<char *>(&len), 4); camline.resize(len); f.read(&camline[0], len); }
            std::istringstream ss(camline);
            uint32_t cw, ch;
            Camera cam = read_camera(ss, cw, ch);
            float bg[3];
            f.read(reinterpret_cast<char *>(bg), sizeof(bg));
            auto tr = read_vec<float>(f, (size_t)n * 10), sh = read_vec<float>(f, (size_t)n * k * 3), op = read_vec<float>(f, n);
            auto skip = read_vec<float>(f, (size_t)w * h * 4);
            auto gt = read_vec<uint32_t>(f, (size_t)w * h);
            const uint32_t cap = 2 * n;
            Context ctx(0, cap, w, h);
            Splats splats(tr.data(), sh.data(), op.data(), n, k);
            DeviceBuffer<uint32_t> d_gt(gt.size());
            d_gt.upload(gt.data(), gt.size());
            TrainConfig cfg;
            cfg.total_train_iters = 1000;
            cfg.seed = 7;
            cfg.max_splats = cap;
            const BoundingBox bounds = bounds_from_pos(ctx, nullptr, BOUND_PERCENTILE, splats.transforms.data(), n);
            std::printf("bounds %.9g %.9g %.9g %.9g %.9g %.9g\n", bounds.center[0], bounds.center[1], bounds.center[2], bounds.extent[0],
                        bounds.extent[1], bounds.extent[2]);
            SplatTrainer trainer(cfg, n, k, bounds);
            const int steps = std::atoi(argv[3]);
            auto one_step = [&]() {
                const float *loss_dev = trainer.step_views(ctx, nullptr, nullptr, {cam}, {d_gt.data()}, w, h, splats);
                float loss;
                check_cuda(cudaMemcpy(&loss, loss_dev, 4, cudaMemcpyDeviceToHost), "loss readback");
                std::printf("loss %.9g\n", loss);
            };
            for (int i = 0; i < steps; i++) one_step();
            const RefineStats rs = trainer.refine(ctx, nullptr, (uint32_t)steps, splats);
            std::printf("refine added %u oversized %u high_grad %u pruned %u non_finite %u total %u\n", rs.num_added, rs.num_split_oversized,
                        rs.num_split_high_grad, rs.num_pruned, rs.num_pruned_non_finite, rs.total_splats);
            one_step();
            return splats.num_splats() == rs.total_splats && trainer.num_splats() == rs.total_splats ? 0 : 1;
        }
    } catch (const Error &e) {
        std::fprintf(stderr, "brush_b200::Error %d: %s\n", e.status, e.what());
        return 2;
    }
    std::fprintf(stderr, "usage: host_check uniforms|fov|errors|render IN OUT\n");
    return 64;
}
