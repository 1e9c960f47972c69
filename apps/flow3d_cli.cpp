// flow3d_cli.cpp -- command-line front end for the solver (the reference's
// main.cpp is a lab script with hard-coded dims and paths; this is the real CLI).
//
//   flow3d_cli --dims W H D --frame0 a.raw --frame1 b.raw [--f32] [--out prefix] [--vtk file.vtk]
//              [--param key=value ...] [--reps N] [--device k | --gpus N] [--verbose]
//              [--diagnostics] [--tolerance T] [--warped PREFIX]
//   flow3d_cli --pairs "frames_%04d.raw" FIRST LAST ...   consecutive frame pairs, solver kept alive
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <future>
#include <string>
#include <vector>

#include "flow3d/cuda_utils.h"
#include "flow3d/data3d.h"
#include "flow3d/optical_flow_e.h"
#include "flow3d_c.h"

namespace {
void usage(const char* a0) {
  std::printf(
      "usage: %s --dims W H D --frame0 F0 --frame1 F1 [--f32] [--out PREFIX] [--vtk FILE]\n"
      "          [--param key=value]... [--reps N] [--device K | --gpus N] [--verbose]\n"
      "          [--diagnostics] [--tolerance T] [--warped PREFIX]\n"
      "       %s --dims W H D --pairs PATTERN FIRST LAST [--f32] [--out PREFIX] ...\n"
      "--gpus N z-shards the volume over devices 0..N-1 (NCCL halo exchange; same bits as one GPU).\n"
      "--diagnostics prints the Jacobi update norm per level; --tolerance T stops a level once the RMS\n"
      "update falls below T voxels (not the reference's fixed iteration count); --warped writes the\n"
      "registered frame 1 (PREFIX_warped.raw) and |warped - frame0| (PREFIX_error.raw), float32.\n"
      "inputs are headerless RAW volumes, x fastest: uint8 by default, float32 with --f32.\n"
      "parameters: warp_levels_count warp_scale_factor outer_iterations_count inner_iterations_count\n"
      "            equation_alpha equation_smoothness equation_data median_radius gaussian_sigma\n",
      a0, a0);
}
bool set_param(flow3d_params& p, const std::string& kv) {
  const size_t eq = kv.find('=');
  if (eq == std::string::npos) return false;
  const std::string k = kv.substr(0, eq), v = kv.substr(eq + 1);
  if (k == "warp_levels_count") p.warp_levels_count = std::strtoull(v.c_str(), nullptr, 10);
  else if (k == "warp_scale_factor") p.warp_scale_factor = std::strtof(v.c_str(), nullptr);
  else if (k == "outer_iterations_count") p.outer_iterations_count = std::strtoull(v.c_str(), nullptr, 10);
  else if (k == "inner_iterations_count") p.inner_iterations_count = std::strtoull(v.c_str(), nullptr, 10);
  else if (k == "equation_alpha") p.equation_alpha = std::strtof(v.c_str(), nullptr);
  else if (k == "equation_smoothness") p.equation_smoothness = std::strtof(v.c_str(), nullptr);
  else if (k == "equation_data") p.equation_data = std::strtof(v.c_str(), nullptr);
  else if (k == "median_radius") p.median_radius = std::strtoull(v.c_str(), nullptr, 10);
  else if (k == "gaussian_sigma") p.gaussian_sigma = std::strtof(v.c_str(), nullptr);
  else return false;
  return true;
}
bool read_frame(Data3D& d, const std::string& path, bool f32, size_t W, size_t H, size_t D) {
  return f32 ? d.ReadRAWFromFileF32(path.c_str(), W, H, D) : d.ReadRAWFromFileU8(path.c_str(), W, H, D);
}
}  // namespace

int main(int argc, char** argv) {
  size_t W = 0, H = 0, D = 0;
  std::string f0, f1, out, vtk, pattern, warped_prefix;
  bool diagnostics = false;
  float tolerance = 0.f;
  long first = 0, last = -1;
  bool f32 = false, verbose = false;
  int reps = 1, device = 0, gpus = 1;
  flow3d_params p;
  flow3d_default_params(&p);
  for (int i = 1; i < argc; ++i) {
    const std::string a = argv[i];
    auto need = [&](int n) { return i + n < argc; };
    if (a == "--dims" && need(3)) { W = std::strtoull(argv[++i], nullptr, 10); H = std::strtoull(argv[++i], nullptr, 10); D = std::strtoull(argv[++i], nullptr, 10); }
    else if (a == "--frame0" && need(1)) f0 = argv[++i];
    else if (a == "--frame1" && need(1)) f1 = argv[++i];
    else if (a == "--pairs" && need(3)) { pattern = argv[++i]; first = std::atol(argv[++i]); last = std::atol(argv[++i]); }
    else if (a == "--out" && need(1)) out = argv[++i];
    else if (a == "--vtk" && need(1)) vtk = argv[++i];
    else if (a == "--f32") f32 = true;
    else if (a == "--diagnostics") diagnostics = true;
    else if (a == "--tolerance" && need(1)) { tolerance = std::strtof(argv[++i], nullptr); diagnostics = true; }
    else if (a == "--warped" && need(1)) warped_prefix = argv[++i];
    else if (a == "--verbose") verbose = true;
    else if (a == "--reps" && need(1)) reps = std::atoi(argv[++i]);
    else if (a == "--device" && need(1)) device = std::atoi(argv[++i]);
    else if (a == "--gpus" && need(1)) gpus = std::atoi(argv[++i]);
    else if (a == "--param" && need(1)) { if (!set_param(p, argv[++i])) { std::printf("bad --param %s\n", argv[i]); return 2; } }
    else { usage(argv[0]); return 2; }
  }
  if (!W || !H || !D || (pattern.empty() && (f0.empty() || f1.empty()))) { usage(argv[0]); return 2; }

  OperationParameters params;
  params.PushValuePtr("warp_levels_count", &p.warp_levels_count);
  params.PushValuePtr("warp_scale_factor", &p.warp_scale_factor);
  params.PushValuePtr("outer_iterations_count", &p.outer_iterations_count);
  params.PushValuePtr("inner_iterations_count", &p.inner_iterations_count);
  params.PushValuePtr("equation_alpha", &p.equation_alpha);
  params.PushValuePtr("equation_smoothness", &p.equation_smoothness);
  params.PushValuePtr("equation_data", &p.equation_data);
  params.PushValuePtr("median_radius", &p.median_radius);
  params.PushValuePtr("gaussian_sigma", &p.gaussian_sigma);

  CUcontext ctx;
  if (!InitCudaContextWithFirstAvailableDevice(&ctx)) return 1;
  OpticalFlowE solver;
  solver.SetDevice(device);
  if (gpus > 1) {
    std::vector<int> devs;
    for (int d = 0; d < gpus; ++d) devs.push_back(d);
    solver.SetDevices(devs);
    if (diagnostics) { std::printf("--diagnostics is single-GPU only\n"); return 2; }
  }
  solver.silent = !verbose;
  DataSize4 size = {W, H, D, 0};
  if (!solver.Initialize(size)) return 1;
  if (diagnostics && !solver.SetDiagnostics(true, tolerance)) return 1;
  Data3D u(W, H, D), v(W, H, D), w(W, H, D);
  Data3D a, b, ahead;
  // pair mode: the file of the NEXT pair's second frame is read by a helper thread while the current pair is
  // being solved (the reference re-reads both frames and re-initialises the solver per pair, main.cpp:132-182)
  std::future<bool> ahead_ready;
  std::string ahead_path;

  std::string next_path;  // set by the pair loop: file to read ahead during this solve ("" = none)
  auto solve_pair = [&](const std::string& p0, const std::string& p1, const std::string& prefix) -> int {
    if (a.Width() == 0 || !pattern.empty()) {
      // in pair mode frame i+1 of the previous pair becomes frame i of this one (no re-read)
      if (!pattern.empty() && b.Width() == W) a.Swap(b);
      else if (!read_frame(a, p0, f32, W, H, D)) return 2;
    }
    if (ahead_ready.valid() && ahead_path == p1) {  // read ahead during the previous solve
      if (!ahead_ready.get()) return 2;
      b.Swap(ahead);
    } else {
      if (ahead_ready.valid()) ahead_ready.get();
      if (!read_frame(b, p1, f32, W, H, D)) return 2;
    }
    if (!next_path.empty()) {
      ahead_path = next_path;
      ahead_ready = std::async(std::launch::async, [&, path = next_path] { return read_frame(ahead, path, f32, W, H, D); });
    }
    for (int r = 0; r < reps; ++r) {
      solver.ComputeFlow(a, b, u, v, w, params);
      if (solver.last_status() != FLOW3D_OK) return 1;
      std::printf("FLOW3D_SOLVE rep=%d total_ms=%.3f device_ms=%.3f mvoxel_per_s=%.3f\n", r, solver.last_total_ms(),
                  solver.last_device_ms(), (double)W * H * D / (solver.last_total_ms() * 1e3));
    }
    if (!prefix.empty()) {
      if (!u.WriteRAWToFileF32((prefix + "_u.raw").c_str()) || !v.WriteRAWToFileF32((prefix + "_v.raw").c_str()) ||
          !w.WriteRAWToFileF32((prefix + "_w.raw").c_str()))
        return 3;
    }
    if (!vtk.empty() && !Data3D::WriteFlowToFileVTK(vtk.c_str(), u, v, w)) return 3;
    if (diagnostics) solver.PrintDiagnostics();
    if (!warped_prefix.empty()) {
      Data3D warped(W, H, D), err(W, H, D);
      if (!solver.WarpFrame(a, b, u, v, w, warped, &err)) return 1;
      const std::string tag = prefix.empty() ? warped_prefix : warped_prefix + prefix.substr(out.size());
      if (!warped.WriteRAWToFileF32((tag + "_warped.raw").c_str()) || !err.WriteRAWToFileF32((tag + "_error.raw").c_str()))
        return 3;
    }
    return 0;
  };

  int rc = 0;
  if (pattern.empty()) {
    rc = solve_pair(f0, f1, out);
  } else {
    char n0[4096], n1[4096];
    for (long i = first; i < last && rc == 0; ++i) {
      std::snprintf(n0, sizeof(n0), pattern.c_str(), (int)i);
      std::snprintf(n1, sizeof(n1), pattern.c_str(), (int)(i + 1));
      next_path.clear();
      if (i + 2 <= last) {
        char n2[4096];
        std::snprintf(n2, sizeof(n2), pattern.c_str(), (int)(i + 2));
        next_path = n2;
      }
      rc = solve_pair(n0, n1, out.empty() ? std::string() : out + "_" + std::to_string(i));
    }
    if (ahead_ready.valid()) ahead_ready.get();
  }
  solver.Destroy();
  return rc;
}
