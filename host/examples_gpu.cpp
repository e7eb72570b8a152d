// examples-gpu -- the reference's CLI driver (haskell/src/test/Examples.hs:26-40) over libpzgpu.so.
//
//   examples-gpu mtt [N]            runMTTExample        (Examples.hs:23-24, MultiTargetTracking.hs:141-145)
//   examples-gpu stt [delay] [N]    STT.runExample       (SingleTargetTracking.hs:61-69); delay=True (the default,
//                                   Examples.hs:31) is the delayed-sampling stream: every particle is the exact
//                                   Gaussian marginal, printed as {"mean":..,"cov":..}; delay=False the bootstrap filter
//   examples-gpu rw1d [N]           runKalman1DParticles (Examples/Demo.hs:200-210), one flat array per line
//                                   for Generic1D.java
//   examples-gpu walker [N]         Walker.run (Examples/Walker.hs:112-148): 1000 measurements of a simulated walker, the
//                                   MStream filter `particles 1000`, one `show`n summary per step:
//                                   ([(frequency,MotionType)..],(mean x,mean y),(mean vx,mean vy))
// Options: --steps T (default: run until the pipe closes, as ZS.runStream does), --emit P (particles
// written per line, default 100: the visualiser draws every particle it is given), --seed S.
//
// Each step simulates ground truth + observations from the generative model on the host (as the
// reference's `generateGroundTruth` / `simulate` do), calls pz_filter_step, and writes ONE JSON line
//   [ [[id,[x..]]..], [ [[id,{"mean":[..],"cov":[[..]..]}]..] per particle ], [[ox,oy(,oz)]..] ]
// on stdout -- the format java/src/main/java/glimpsetracks/App.java:347-399 parses, so
//   examples-gpu mtt 1000000 | mvn exec:java -Dexec.mainClass=glimpsetracks.App
// works as the reference's README describes.  This stands in for the Haskell host (haskell/PzGpu.hs),
// which cannot be compiled in this image.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <vector>

#include "pzgpu.h"

namespace {

void die(const char* what) {
  fprintf(stderr, "examples-gpu: %s: %s\n", what, pz_last_error());
  exit(1);
}

void put_vec(std::string& s, const double* v, int n) {
  char buf[40];
  s += '[';
  for (int i = 0; i < n; ++i) {
    if (std::isfinite(v[i])) snprintf(buf, sizeof(buf), "%.9g", v[i]);
    else snprintf(buf, sizeof(buf), "null");  // aeson encodes NaN / Infinity as null
    if (i) s += ',';
    s += buf;
  }
  s += ']';
}

struct Sim {
  std::mt19937_64 gen;
  std::normal_distribution<double> nd{0.0, 1.0};
  std::uniform_real_distribution<double> ud{0.0, 1.0};
  explicit Sim(uint64_t seed) : gen(seed) {}
  double n() { return nd(gen); }
  double u() { return ud(gen); }
  int poisson(double lambda) { return std::poisson_distribution<int>(lambda)(gen); }
};

// x' ~ N(F x, Q) with diagonal Q (every in-scope model), y ~ N(H x', R) with diagonal R
void lg_sim_step(const pz_model_desc& d, Sim& sim, std::vector<double>& x, std::vector<double>& y) {
  const int nx = d.nx, ny = d.ny;
  std::vector<double> xn(nx);
  for (int i = 0; i < nx; ++i) {
    double s = 0;
    for (int k = 0; k < nx; ++k) s += d.F[i * nx + k] * x[k];
    xn[i] = s + std::sqrt(d.Q[i * nx + i]) * sim.n();
  }
  x = xn;
  y.assign(ny, 0.0);
  for (int i = 0; i < ny; ++i) {
    double s = 0;
    for (int k = 0; k < nx; ++k) s += d.H[i * nx + k] * x[k];
    y[i] = s + std::sqrt(d.R[i * ny + i]) * sim.n();
  }
}

int run_lg(const pz_model_desc& d, uint64_t n, long steps, int emit, uint64_t seed, bool flat) {
  const bool delayed = d.delayed != 0;
  std::vector<double> pmean(d.nx), pcov((size_t)d.nx * d.nx);
  pz_filter* f = nullptr;
  if (pz_filter_create(&d, n, seed, 0, &f)) die("pz_filter_create");
  Sim sim(seed ^ 0x9E3779B97F4A7C15ull);
  std::vector<double> x(d.nx, 0.0), y;
  const int32_t stride = (int32_t)((n + emit - 1) / emit);
  const int64_t n_sel = (int64_t)((n + stride - 1) / stride);
  std::vector<double> parts((size_t)n_sel * d.nx);
  std::string line;
  for (long t = 0; steps < 0 || t < steps; ++t) {
    lg_sim_step(d, sim, x, y);
    pz_step_summary s;
    if (pz_filter_step(f, y.data(), 1, d.ny, &s)) die("pz_filter_step");
    line.clear();
    if (flat) {  // (observedPosition, actualPosition, average particles), Demo.hs:207
      double v[3] = {y[0], x[0], s.mean[0]};
      put_vec(line, v, 3);
    } else if (delayed) {
      // every particle is the symbolic marginal: ToJSON MDistr, DelayedSampling.hs:69-73
      if (pz_filter_read_posterior(f, pmean.data(), pcov.data())) die("pz_filter_read_posterior");
      std::string node = "[[0,{\"mean\":"; put_vec(node, pmean.data(), d.nx); node += ",\"cov\":[";
      for (int r = 0; r < d.nx; ++r) { if (r) node += ','; put_vec(node, &pcov[(size_t)r * d.nx], d.nx); }
      node += "]}]]";
      line += "[[[0,"; put_vec(line, x.data(), d.nx); line += "]],[";
      for (int64_t i = 0; i < n_sel; ++i) { if (i) line += ','; line += node; }
      line += "],["; put_vec(line, y.data(), d.ny); line += "]]";
    } else {
      if (pz_filter_read_particles(f, stride, parts.data(), (int64_t)parts.size())) die("pz_filter_read_particles");
      line += "[[[0,"; put_vec(line, x.data(), d.nx); line += "]],[";   // [(0, groundTruth)]
      for (int64_t i = 0; i < n_sel; ++i) {                             // map (\x -> [(0, x)]) particles; RConst -> bare array
        if (i) line += ',';
        line += "[[0,"; put_vec(line, &parts[(size_t)i * d.nx], d.nx); line += "]]";
      }
      line += "],["; put_vec(line, y.data(), d.ny); line += "]]";
    }
    if (puts(line.c_str()) < 0 || fflush(stdout) != 0) break;  // the consumer closed the pipe
  }
  pz_filter_destroy(f);
  return 0;
}

// ---- Walker.run ----------------------------------------------------------------------------------------------------
// Haskell's `show` of a Double: shortest digits that round-trip; decimal notation for 0.1 <= |x| < 10^7, else d.ddde<n>
std::string show_double(double x) {
  if (std::isnan(x)) return "NaN";
  if (std::isinf(x)) return x > 0 ? "Infinity" : "-Infinity";
  char buf[64];
  int prec = 1;
  for (; prec <= 17; ++prec) { snprintf(buf, sizeof(buf), "%.*e", prec - 1, x); if (strtod(buf, nullptr) == x) break; }
  std::string m(buf);                       // [-]d.ddde[+-]xx
  const bool neg = m[0] == '-';
  if (neg) m.erase(0, 1);
  const size_t epos = m.find('e');
  const int ex = atoi(m.c_str() + epos + 1);
  std::string digits;
  for (size_t i = 0; i < epos; ++i) if (m[i] != '.') digits += m[i];
  std::string out;
  if (x == 0) out = "0.0";
  else if (ex >= -1 && ex < 7) {
    if (ex == -1) out = "0." + digits;
    else {
      std::string ip = digits.substr(0, std::min<size_t>(digits.size(), (size_t)ex + 1));
      while ((int)ip.size() < ex + 1) ip += '0';
      std::string fp = digits.size() > (size_t)ex + 1 ? digits.substr(ex + 1) : "0";
      out = ip + "." + fp;
    }
  } else {
    out = digits.substr(0, 1) + "." + (digits.size() > 1 ? digits.substr(1) : "0") + "e" + std::to_string(ex);
  }
  return (neg ? "-" : "") + out;
}

struct WalkerTruth { double x = 0, y = 0, vx = 0, vy = 0; int mt = 0; };

void walker_init_velocity(const pz_walker_params& w, Sim& sim, int mt, double& vx, double& vy) {   // initVelocity, Walker.hs:51-64
  if (w.speed_hi[mt] <= 0) { vx = vy = 0; return; }
  const double speed = w.speed_lo[mt] + (w.speed_hi[mt] - w.speed_lo[mt]) * sim.u();
  const double ang = 2 * M_PI * sim.u();
  vx = speed * std::cos(ang); vy = speed * std::sin(ang);
}
void walker_coast(const pz_walker_params& w, Sim& sim, double dt, WalkerTruth& s) {   // coast, Walker.hs:31-39 (D.normal takes a VARIANCE)
  const double sd = std::sqrt(std::sqrt(dt) * w.pos_noise);
  s.x += s.vx * dt + sd * sim.n();
  s.y += s.vy * dt + sd * sim.n();
}
void walker_motion(const pz_walker_params& w, Sim& sim, double dt, WalkerTruth& s) {   // motion, Walker.hs:66-80
  for (int guard = 0; guard < 100000; ++guard) {
    double rate[3], tot = 0;
    for (int to = 0; to < 3; ++to) { rate[to] = w.half_life[s.mt][to] > 0 ? std::log(2.0) / w.half_life[s.mt][to] : 0; tot += rate[to]; }
    const double e = -std::log(1.0 - sim.u());
    const double t_tr = w.exp_mean_as_written ? e * tot : e / tot;   // D.exponential is handed recip(sum rates) as its RATE (:69)
    if (t_tr > dt) { walker_coast(w, sim, dt, s); return; }
    walker_coast(w, sim, t_tr, s);
    double u = sim.u() * tot; int mt = 2;
    for (int to = 0; to < 3; ++to) { if (u < rate[to]) { mt = to; break; } u -= rate[to]; }
    s.mt = mt;
    walker_init_velocity(w, sim, mt, s.vx, s.vy);
    dt -= t_tr;
  }
}

int run_walker(uint64_t n, long steps, uint64_t seed) {
  pz_model_desc d;
  if (pz_model_walker(&d)) die("pz_model_walker");
  pz_filter* f = nullptr;
  if (pz_filter_create(&d, n, seed, 0, &f)) die("pz_filter_create");
  const pz_walker_params& w = d.walker;
  Sim sim(seed ^ 0xA0761D6478BD642Full);
  WalkerTruth tr;   // walkerInit (:105-111)
  { double u = sim.u(); tr.mt = u < w.p_init[0] ? 0 : (u < w.p_init[0] + w.p_init[1] ? 1 : 2); walker_init_velocity(w, sim, tr.mt, tr.vx, tr.vy); }
  if (steps < 0) steps = 1000;   // generateWalker 1000 (:147)
  std::vector<double> parts((size_t)n * 5);
  static const char* names[3] = {"Stationary", "Walking", "Running"};
  for (long t = 0; t < steps; ++t) {
    walker_motion(w, sim, w.dt, tr);
    const double y[2] = {tr.x + w.obs_std * sim.n(), tr.y + w.obs_std * sim.n()};   // walkerGenMeasurement (:96-103)
    pz_step_summary s;
    if (pz_filter_step(f, y, 1, 2, &s)) die("pz_filter_step");
    if (pz_filter_read_particles(f, 1, parts.data(), (int64_t)parts.size())) die("pz_filter_read_particles");
    double cnt[3] = {0, 0, 0}, m[4] = {0, 0, 0, 0};
    for (uint64_t i = 0; i < n; ++i) {
      const double* p = &parts[(size_t)i * 5];
      for (int k = 0; k < 4; ++k) m[k] += p[k];
      const int mt = (int)p[4];
      if (mt >= 0 && mt < 3) cnt[mt] += 1;
    }
    std::string line = "([";   // summarize (:137-143): frequencies in constructor order, then the two averaged pairs
    bool first = true;
    for (int k = 0; k < 3; ++k)
      if (cnt[k] > 0) { line += (first ? "(" : ",(") + show_double(cnt[k] / (double)n) + "," + names[k] + ")"; first = false; }
    line += "],(" + show_double(m[0] / n) + "," + show_double(m[1] / n) + "),(" + show_double(m[2] / n) + "," + show_double(m[3] / n) + "))";
    if (puts(line.c_str()) < 0 || fflush(stdout) != 0) break;
  }
  pz_filter_destroy(f);
  return 0;
}

struct GtTrack { int id; double x[4]; };

int run_mtt(uint64_t n, long steps, int emit, uint64_t seed) {
  pz_model_desc d;
  if (pz_model_mtt(&d)) die("pz_model_mtt");
  pz_filter* f = nullptr;
  if (pz_filter_create(&d, n, seed, 0, &f)) die("pz_filter_create");
  Sim sim(seed ^ 0xD1B54A32D192ED03ull);
  const double pspd = d.mtt.survival_prob * d.mtt.pd, p_coast = (1 - d.mtt.pd) / (1 - pspd);
  std::vector<GtTrack> tracks;
  int next_id = 0;
  const int32_t stride = (int32_t)((n + emit - 1) / emit);
  const int64_t n_sel = (int64_t)((n + stride - 1) / stride);
  std::vector<int32_t> cnt(n_sel), ids(n_sel * 16);
  std::vector<double> mean(n_sel * 16 * 4), cov(n_sel * 16 * 16);
  std::string line;
  for (long t = 0; steps < 0 || t < steps; ++t) {
    // sampleStep under MP.sim (MultiTargetTracking.hs:89-99,126-131)
    for (auto& tr : tracks) {
      double xn[4];
      for (int i = 0; i < 4; ++i) {
        double s = 0;
        for (int k = 0; k < 4; ++k) s += d.F[i * 4 + k] * tr.x[k];
        xn[i] = s + std::sqrt(d.Q[i * 4 + i]) * sim.n();
      }
      memcpy(tr.x, xn, sizeof(xn));
    }
    std::vector<int> keys;
    const int n_c = sim.poisson(d.mtt.clutter_lambda), n_n = sim.poisson(d.mtt.new_track_lambda);
    for (int i = 0; i < n_c; ++i) keys.push_back(-1);
    for (int i = 0; i < n_n; ++i) keys.push_back(-2);
    std::vector<char> det(tracks.size(), 0);
    for (size_t k = 0; k < tracks.size(); ++k) if (sim.u() < pspd) { det[k] = 1; keys.push_back((int)k); }
    for (size_t i = keys.size(); i > 1; --i) std::swap(keys[i - 1], keys[(size_t)(sim.u() * i) % i]);  // randomlyInterleave
    std::vector<double> obs;
    std::vector<GtTrack> born;
    for (int key : keys) {
      if ((int)obs.size() / 2 >= d.mtt.m_max) break;
      double z0, z1;
      if (key == -1) { z0 = std::sqrt(d.mtt.clutter_var) * sim.n(); z1 = std::sqrt(d.mtt.clutter_var) * sim.n(); }
      else if (key == -2) {
        GtTrack nt; nt.id = next_id++;
        nt.x[0] = std::sqrt(d.mtt.birth_pos_var) * sim.n(); nt.x[1] = std::sqrt(d.mtt.birth_pos_var) * sim.n();
        nt.x[2] = std::sqrt(d.mtt.birth_vel_var) * sim.n(); nt.x[3] = std::sqrt(d.mtt.birth_vel_var) * sim.n();
        born.push_back(nt);
        z0 = nt.x[0] + std::sqrt(d.mtt.meas_var) * sim.n(); z1 = nt.x[1] + std::sqrt(d.mtt.meas_var) * sim.n();
      } else { z0 = tracks[key].x[0] + std::sqrt(d.mtt.meas_var) * sim.n(); z1 = tracks[key].x[1] + std::sqrt(d.mtt.meas_var) * sim.n(); }
      obs.push_back(z0); obs.push_back(z1);
    }
    std::vector<GtTrack> next;
    for (size_t k = 0; k < tracks.size(); ++k) if (det[k] || sim.u() < p_coast) next.push_back(tracks[k]);
    for (auto& b : born) if ((int)next.size() < d.mtt.k_max) next.push_back(b);
    tracks.swap(next);

    pz_step_summary s;
    if (pz_filter_step(f, obs.data(), (int32_t)(obs.size() / 2), 2, &s)) die("pz_filter_step");
    if (pz_filter_read_tracks(f, stride, cnt.data(), ids.data(), mean.data(), cov.data(), n_sel)) die("pz_filter_read_tracks");
    line.clear();
    line += "[[";
    for (size_t k = 0; k < tracks.size(); ++k) {
      if (k) line += ',';
      line += '[' + std::to_string(tracks[k].id) + ','; put_vec(line, tracks[k].x, 4); line += ']';
    }
    line += "],[";
    for (int64_t i = 0; i < n_sel; ++i) {
      if (i) line += ',';
      line += '[';
      for (int k = 0; k < cnt[i]; ++k) {
        if (k) line += ',';
        line += '[' + std::to_string(ids[i * 16 + k]) + ",{\"mean\":"; put_vec(line, &mean[(i * 16 + k) * 4], 4);
        line += ",\"cov\":[";
        for (int r = 0; r < 4; ++r) { if (r) line += ','; put_vec(line, &cov[(i * 16 + k) * 16 + r * 4], 4); }
        line += "]}]";
      }
      line += ']';
    }
    line += "],[";
    for (size_t o = 0; o < obs.size() / 2; ++o) { if (o) line += ','; put_vec(line, &obs[2 * o], 2); }
    line += "]]";
    if (puts(line.c_str()) < 0 || fflush(stdout) != 0) break;
  }
  pz_filter_destroy(f);
  return 0;
}

void usage() {
  fputs("usage: examples-gpu {mtt [N] | stt [delay] [N] | rw1d [N] | walker [N]} [--steps T] [--emit P] [--seed S]\n"
        "  one JSON line per time step on stdout (java/src/main/java/glimpsetracks/App.java, Generic1D.java)\n", stderr);
}

}  // namespace

int main(int argc, char** argv) {
  std::vector<std::string> pos;
  long steps = -1; int emit = 100; uint64_t seed = 1;
  for (int i = 1; i < argc; ++i) {
    std::string a = argv[i];
    if (a == "--steps" && i + 1 < argc) steps = atol(argv[++i]);
    else if (a == "--emit" && i + 1 < argc) emit = atoi(argv[++i]);
    else if (a == "--seed" && i + 1 < argc) seed = strtoull(argv[++i], nullptr, 10);
    else if (a == "-h" || a == "--help") { usage(); return 0; }
    else pos.push_back(a);
  }
  if (pos.empty() || emit < 1) { usage(); return 2; }
  auto nth = [&](size_t i, const char* dflt) { return i < pos.size() ? pos[i] : std::string(dflt); };
  if (pos[0] == "mtt") return run_mtt(strtoull(nth(1, "100").c_str(), nullptr, 10), steps, emit, seed);  // default N = 100, Examples.hs:30
  if (pos[0] == "stt") {
    const std::string delay = nth(1, "True");                                                         // Examples.hs:31
    if (delay != "True" && delay != "False") { usage(); return 2; }                                   // `read` of a Bool
    pz_model_desc d;
    if (pz_model_stt(&d)) die("pz_model_stt");
    d.delayed = delay == "True";
    return run_lg(d, strtoull(nth(2, "100").c_str(), nullptr, 10), steps, emit, seed, false);
  }
  if (pos[0] == "rw1d") {
    pz_model_desc d;
    if (pz_model_rw1d(&d, 1.0, 3.0)) die("pz_model_rw1d");
    return run_lg(d, strtoull(nth(1, "1000").c_str(), nullptr, 10), steps, emit, seed, true);             // Demo.hs:206
  }
  if (pos[0] == "walker") return run_walker(strtoull(nth(1, "1000").c_str(), nullptr, 10), steps, seed);   // particles 1000, Walker.hs:135
  usage();
  return 2;
}
