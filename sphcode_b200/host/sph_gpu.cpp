// sph_gpu — the reference's command line on the device path:   sph_gpu <sample | parameter.json> [threads]
//
// Replaces, for the hot path's callers and data formats, the Boost-dependent parts of
// mitchiinaga/sphcode that sit either side of the per-step path:
//   * Solver::read_parameterfile (src/solver.cpp:155-299): same JSON keys, defaults and error texts,
//     read by a small JSON reader of our own (the reference needs boost::property_tree for this);
//     the sample names resolve to the shipped sample/<name>/<name>.json if it exists under the
//     current directory, else to the shipped values restated below;
//   * Solver::make_initial_condition + src/sample/*.cpp: the six generators, DIM at run time, the
//     3-D lattice of evrard filled by all host threads (the reference's emplace_back loop is serial);
//   * Solver::run (src/solver.cpp:301-350): initialize, then integrate until endTime, snapshots every
//     outputTime and energies every energyTime in the reference's text formats (src/output.cpp:14-90);
//     the step itself is sphb_integrate of libsphb.so, the state stays in HBM between outputs.
// Extensions (do not exist in the reference): `--set key=value` overrides a JSON key (e.g. N,
// SPHType) without editing the file, `--no-snapshots` keeps only energy.dat, `--binary-snapshots` writes
// NNNNN.bin (full-precision SPHParticle records behind a 32-byte header {"SPHB", dim, n, record bytes, time})
// next to the 6-digit text files, `--steps n` stops after n steps, `--dump-ic file` writes the initial
// SPHParticle array (binary) and `--dump-params file` the resolved parameters (text); both exit without touching a GPU.
// "threads" is accepted and only used for the host-side generators.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <fstream>
#include <iostream>
#include <map>
#include <random>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>
#include <sys/stat.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../../include/sphb.h"

namespace {

// ---------------------------------------------------------------------------------------------
// JSON: objects of strings / numbers / booleans / arrays of numbers (all the parameter files use)
// ---------------------------------------------------------------------------------------------
struct JValue {
    enum Kind { NUL, BOOL, NUM, STR, ARR } kind = NUL;
    bool b = false;
    double num = 0.0;
    std::string str;            // STR: the text; NUM: the literal as written
    std::vector<double> arr;
};
using JObject = std::map<std::string, JValue>;

struct JParser {
    const std::string & s;
    size_t i = 0;
    explicit JParser(const std::string & text) : s(text) {}
    [[noreturn]] void fail(const char * what) const
    {
        std::ostringstream o;
        o << "json: " << what << " at offset " << i;
        throw std::runtime_error(o.str());
    }
    void ws() { while (i < s.size() && std::isspace((unsigned char)s[i])) ++i; }
    bool eat(char c) { ws(); if (i < s.size() && s[i] == c) { ++i; return true; } return false; }
    std::string string_()
    {
        ws();
        if (i >= s.size() || s[i] != '"') fail("expected string");
        ++i;
        std::string out;
        while (i < s.size() && s[i] != '"') {
            if (s[i] == '\\' && i + 1 < s.size()) {
                const char e = s[++i];
                out += e == 'n' ? '\n' : e == 't' ? '\t' : e;
            } else out += s[i];
            ++i;
        }
        if (i >= s.size()) fail("unterminated string");
        ++i;
        return out;
    }
    double number_(std::string * lit = nullptr)
    {
        ws();
        const size_t b = i;
        while (i < s.size() && (std::isdigit((unsigned char)s[i]) || std::strchr("+-.eE", s[i]))) ++i;
        if (b == i) fail("expected number");
        const std::string t = s.substr(b, i - b);
        if (lit) *lit = t;
        return std::stod(t);            // the reference parses through std::stod as well (src/solver.cpp:269)
    }
    JValue value_()
    {
        ws();
        JValue v;
        if (i >= s.size()) fail("unexpected end");
        const char c = s[i];
        if (c == '"') { v.kind = JValue::STR; v.str = string_(); }
        else if (c == '[') {
            ++i;
            v.kind = JValue::ARR;
            if (!eat(']')) {
                do { v.arr.push_back(number_()); } while (eat(','));
                if (!eat(']')) fail("expected ]");
            }
        } else if (!s.compare(i, 4, "true")) { v.kind = JValue::BOOL; v.b = true; i += 4; }
        else if (!s.compare(i, 5, "false")) { v.kind = JValue::BOOL; v.b = false; i += 5; }
        else if (!s.compare(i, 4, "null")) { i += 4; }
        else { v.kind = JValue::NUM; v.num = number_(&v.str); }
        return v;
    }
    JObject object_()
    {
        JObject o;
        if (!eat('{')) fail("expected {");
        if (eat('}')) return o;
        do {
            const std::string k = string_();
            if (!eat(':')) fail("expected :");
            o[k] = value_();
        } while (eat(','));
        if (!eat('}')) fail("expected }");
        return o;
    }
};

JObject read_json_file(const std::string & path)
{
    std::ifstream f(path);
    if (!f) throw std::runtime_error(path + ": cannot open file");      // boost: "<file>: cannot open file"
    std::stringstream ss;
    ss << f.rdbuf();
    const std::string text = ss.str();
    JParser p(text);
    return p.object_();
}

// `--set key=value`: booleans, numbers, [a,b] arrays, anything else a string
JValue parse_override(const std::string & text)
{
    JValue v;
    if (text == "true" || text == "false") { v.kind = JValue::BOOL; v.b = text == "true"; return v; }
    if (!text.empty() && text[0] == '[') { JParser p(text); return p.value_(); }
    char * end = nullptr;
    const double d = std::strtod(text.c_str(), &end);
    if (end && *end == 0 && !text.empty()) { v.kind = JValue::NUM; v.num = d; v.str = text; return v; }
    v.kind = JValue::STR; v.str = text;
    return v;
}

// ---------------------------------------------------------------------------------------------
// parameters (include/parameters.hpp:20-79 through the JSON keys of src/solver.cpp:155-299)
// ---------------------------------------------------------------------------------------------
struct Sample { const char * name; int dim; int default_n; const char * shipped_json; };

// The shipped sample/<name>/<name>.json files, restated (used when the file is not under the cwd).
const Sample SAMPLES[] = {
    {"shock_tube", 1, 100,
     R"({"outputDirectory":"sample/shock_tube/results","endTime":0.2,"avAlpha":1.0,"neighborNumber":4,"gamma":1.4,
         "kernel":"cubic_spline","N":50,"periodic":true,"iterativeSmoothingLength":true,"rangeMax":[1.5],"rangeMin":[-0.5],
         "useTimeDependentAV":false,"SPHType":"ssph"})"},
    {"gresho_chan_vortex", 2, 64,
     R"({"outputDirectory":"sample/gresho_chan_vortex/results","endTime":1.0,"avAlpha":1.0,"neighborNumber":32,
         "useBalsaraSwitch":true,"leafParticleNumber":32,"gamma":1.66666666666666666666666666666666667,"kernel":"wendland",
         "N":64,"periodic":true,"rangeMax":[0.5,0.5],"rangeMin":[-0.5,-0.5],"SPHType":"ssph"})"},
    {"pairing_instability", 2, 64,
     R"({"outputDirectory":"sample/pairing_instability/results","endTime":1.0,"avAlpha":1.0,"neighborNumber":32,
         "leafParticleNumber":32,"gamma":1.66666666666666666666666666666666667,"kernel":"cubic_spline","N":64,
         "periodic":true,"rangeMax":[0.5,0.5],"rangeMin":[-0.5,-0.5]})"},
    {"hydrostatic", 2, 32,
     R"({"outputDirectory":"sample/hydrostatic/results","endTime":8.0,"avAlpha":1.0,"neighborNumber":32,
         "useBalsaraSwitch":true,"leafParticleNumber":16,"gamma":1.66666666666666666666666666666666667,"kernel":"wendland",
         "N":32,"periodic":true,"rangeMax":[0.5,0.5],"rangeMin":[-0.5,-0.5],"SPHType":"disph"})"},
    {"khi", 2, 128,
     R"({"outputDirectory":"sample/khi/results","endTime":3.0,"outputTime":0.1,"avAlpha":1.0,"neighborNumber":32,
         "useBalsaraSwitch":true,"useTimeDependentAV":true,"useArtificialConductivity":false,"leafParticleNumber":16,
         "gamma":1.66666666666666666666666666666666667,"kernel":"wendland","N":256,"periodic":true,
         "rangeMax":[1.0,1.0],"rangeMin":[0.0,0.0],"SPHType":"ssph"})"},
    {"evrard", 3, 20,
     R"({"outputDirectory":"sample/evrard/results","endTime":3.0,"avAlpha":1.0,"neighborNumber":32,"useBalsaraSwitch":true,
         "useTimeDependentAV":true,"useArtificialConductivity":false,"leafParticleNumber":32,
         "gamma":1.66666666666666666666666666666666667,"kernel":"wendland","N":30,"periodic":false,"useGravity":true,
         "SPHType":"disph"})"},
};

const Sample * find_sample(const std::string & name)
{
    for (const Sample & s : SAMPLES) if (name == s.name) return &s;
    return nullptr;
}

struct Run {
    sphb_params p{};
    const Sample * sample = nullptr;
    int n_side = 0;
    std::string output_dir;
    double t_start = 0.0, t_end = 0.0, t_output = 0.0, t_energy = 0.0;
};

struct Getter {
    const JObject & o;
    const JValue * find(const char * k) const { auto it = o.find(k); return it == o.end() ? nullptr : &it->second; }
    [[noreturn]] static void missing(const char * k) { throw std::runtime_error(std::string("No such node (") + k + ")"); }
    double real(const char * k) const { const JValue * v = find(k); if (!v) missing(k); return num(*v, k); }
    double real(const char * k, double d) const { const JValue * v = find(k); return v ? num(*v, k) : d; }
    int integer(const char * k, int d) const { const JValue * v = find(k); return v ? (int)num(*v, k) : d; }
    bool boolean(const char * k, bool d) const
    {
        const JValue * v = find(k);
        if (!v) return d;
        if (v->kind == JValue::BOOL) return v->b;
        if (v->kind == JValue::NUM) return v->num != 0.0;
        if (v->kind == JValue::STR) return v->str == "true" || v->str == "1";
        throw std::runtime_error(std::string("conversion of data to type \"bool\" failed (") + k + ")");
    }
    std::string text(const char * k, const char * d) const
    {
        const JValue * v = find(k);
        if (!v) { if (!d) missing(k); return d; }
        return v->str;
    }
    static double num(const JValue & v, const char * k)
    {
        if (v.kind == JValue::NUM) return v.num;
        if (v.kind == JValue::STR) return std::stod(v.str);
        throw std::runtime_error(std::string("conversion of data to type \"double\" failed (") + k + ")");
    }
};

// Same order, defaults and error texts as Solver::read_parameterfile.
Run resolve(const std::string & arg, const std::vector<std::pair<std::string, std::string>> & overrides)
{
    Run r;
    JObject in;
    r.sample = find_sample(arg);
    if (r.sample) {
        const std::string shipped = std::string("sample/") + arg + "/" + arg + ".json";
        struct stat st;
        if (stat(shipped.c_str(), &st) == 0) in = read_json_file(shipped);
        else { const std::string text = r.sample->shipped_json; JParser p(text); in = p.object_(); }
    } else {
        in = read_json_file(arg);
        auto it = in.find("sample");                         // extension: a parameter file may name its sample
        if (it != in.end()) r.sample = find_sample(it->second.str);
        if (!r.sample) throw std::runtime_error("unknown sample type.");          // src/solver.cpp:491
    }
    for (auto & kv : overrides) in[kv.first] = parse_override(kv.second);
    const Getter g{in};
    const int dim = r.sample->dim;
    r.n_side = g.integer("N", r.sample->default_n);
    r.output_dir = g.text("outputDirectory", nullptr);
    r.t_start = g.real("startTime", 0.0);
    r.t_end = g.real("endTime");
    if (r.t_end < r.t_start) throw std::runtime_error("endTime < startTime");
    r.t_output = g.real("outputTime", (r.t_end - r.t_start) / 100);
    r.t_energy = g.real("energyTime", r.t_output);

    sphb_params & p = r.p;
    const std::string type = g.text("SPHType", "ssph");
    if (type == "ssph") p.sph_type = SPHB_SSPH;
    else if (type == "disph") p.sph_type = SPHB_DISPH;
    else if (type == "gsph") p.sph_type = SPHB_GSPH;
    else throw std::runtime_error("Unknown SPH type");
    p.cfl_sound = g.real("cflSound", 0.3);
    p.cfl_force = g.real("cflForce", 0.125);
    p.av_alpha = g.real("avAlpha", 1.0);
    p.use_balsara_switch = g.boolean("useBalsaraSwitch", true);
    p.use_time_dependent_av = g.boolean("useTimeDependentAV", false);
    p.alpha_max = 2.0; p.alpha_min = 0.1; p.epsilon_av = 0.2;
    if (p.use_time_dependent_av) {
        p.alpha_max = g.real("alphaMax", 2.0);
        p.alpha_min = g.real("alphaMin", 0.1);
        if (p.alpha_max < p.alpha_min) throw std::runtime_error("alphaMax < alphaMin");
        p.epsilon_av = g.real("epsilonAV", 0.2);
    }
    p.use_ac = g.boolean("useArtificialConductivity", false);
    p.alpha_ac = p.use_ac ? g.real("alphaAC", 1.0) : 1.0;
    p.max_tree_level = g.integer("maxTreeLevel", 20);
    p.leaf_particle_num = g.integer("leafParticleNumber", 1);
    p.neighbor_number = g.integer("neighborNumber", 32);
    p.gamma = g.real("gamma");
    const std::string kernel = g.text("kernel", "cubic_spline");
    if (kernel == "cubic_spline") p.kernel = SPHB_CUBIC_SPLINE;
    else if (kernel == "wendland") p.kernel = SPHB_WENDLAND;
    else throw std::runtime_error("kernel is unknown.");
    p.iterative_sml = g.boolean("iterativeSmoothingLength", true);
    p.periodic = g.boolean("periodic", false);
    if (p.periodic) {
        const JValue * mx = g.find("rangeMax"), * mn = g.find("rangeMin");
        if (!mx) Getter::missing("rangeMax");
        if ((int)mx->arr.size() != dim) throw std::runtime_error("rangeMax != DIM");
        if (!mn) Getter::missing("rangeMin");
        if ((int)mn->arr.size() != dim) throw std::runtime_error("rangeMax != DIM");      // sic, src/solver.cpp:277
        for (int d = 0; d < dim; ++d) { p.range_max[d] = mx->arr[d]; p.range_min[d] = mn->arr[d]; }
    }
    p.use_gravity = g.boolean("useGravity", false);
    p.G = 1.0; p.theta = 0.5;
    if (p.use_gravity) { p.G = g.real("G", 1.0); p.theta = g.real("theta", 0.5); }
    p.gsph_2nd_order = p.sph_type == SPHB_GSPH ? g.boolean("use2ndOrderGSPH", true) : 1;
    return r;
}

// ---------------------------------------------------------------------------------------------
// initial conditions (src/sample/*.cpp); records have the SPHParticle layout of the sample's DIM
// ---------------------------------------------------------------------------------------------
struct Particles {
    int dim = 0;
    size_t rec = 0;
    std::vector<unsigned char> bytes;
    size_t size() const { return rec ? bytes.size() / rec : 0; }
    void resize(int d, size_t n) { dim = d; rec = sphb_sizeof_particle(d); bytes.assign(n * rec, 0); }
    double * rec_d(size_t i) { return reinterpret_cast<double *>(bytes.data() + i * rec); }
    const double * rec_d(size_t i) const { return reinterpret_cast<const double *>(bytes.data() + i * rec); }
    // member offsets in doubles (include/particle.hpp:8-33)
    int o_pos() const { return 0; }
    int o_vel() const { return dim; }
    int o_acc() const { return 3 * dim; }
    int o_scalar(int k) const { return 4 * dim + k; }      // 0 mass 1 dens 2 pres 3 ene 4 ene_p 5 dene 6 sml 7 sound 8 balsara 9 alpha 10 gradh 11 phi
    int * rec_i(size_t i) { return reinterpret_cast<int *>(rec_d(i) + 4 * dim + 12); }     // id, neighbor
    const int * rec_i(size_t i) const { return reinterpret_cast<const int *>(rec_d(i) + 4 * dim + 12); }
};
enum { S_MASS = 0, S_DENS, S_PRES, S_ENE, S_ENE_P, S_DENE, S_SML, S_SOUND, S_BALSARA, S_ALPHA, S_GRADH, S_PHI };

void finish_ideal_gas(Particles & q, double gamma)        // ene = P / ((gamma - 1) rho), id = index
{
    const size_t n = q.size();
    for (size_t i = 0; i < n; ++i) {
        double * r = q.rec_d(i);
        r[q.o_scalar(S_ENE)] = r[q.o_scalar(S_PRES)] / ((gamma - 1.0) * r[q.o_scalar(S_DENS)]);
        q.rec_i(i)[0] = (int)i;
    }
}

// src/sample/shock_tube.cpp:18-51: 8N particles at dx/4 left of x = 0.5, 2N at dx right of it
void make_shock_tube(Particles & q, int N, double gamma)
{
    const double dx_r = 0.5 / N, dx_l = dx_r * 0.25;
    const int num = N * 10;
    q.resize(1, num);
    double x = -0.5 + dx_l * 0.5, dx = dx_l, dens = 1.0, pres = 1.0;
    const double mass = 0.5 / N * 0.25;
    bool left = true;
    for (int i = 0; i < num; ++i) {
        double * r = q.rec_d(i);
        r[0] = x;
        r[q.o_scalar(S_DENS)] = dens; r[q.o_scalar(S_PRES)] = pres; r[q.o_scalar(S_MASS)] = mass;
        x += dx;
        if (x > 0.5 && left) { x = 0.5 + dx_r * 0.5; dx = dx_r; dens = 0.25; pres = 0.1795; left = false; }
    }
    finish_ideal_gas(q, gamma);
}

// src/sample/khi.cpp:18-73: two density layers on the unit square, staggered rows in the thin layer
void make_khi(Particles & q, int N, double gamma)
{
    const int num = N * N * 3 / 4;
    const double dx = 1.0 / N, mass = 1.5 / num;
    q.resize(2, num);
    double x = dx * 0.5, y = dx * 0.5;
    int region = 1;
    bool odd = true;
    const double sigma2_inv = 2 / (0.05 * 0.05);
    for (int i = 0; i < num; ++i) {
        double * r = q.rec_d(i);
        r[0] = x; r[1] = y;
        r[q.o_vel()] = region == 1 ? -0.5 : 0.5;
        r[q.o_vel() + 1] = 0.1 * std::sin(4.0 * M_PI * x) * (std::exp(-(y - 0.25) * (y - 0.25) * 0.5 * sigma2_inv)
                                                            + std::exp(-(y - 0.75) * (y - 0.75) * 0.5 * sigma2_inv));
        r[q.o_scalar(S_MASS)] = mass;
        r[q.o_scalar(S_DENS)] = (double)region;
        r[q.o_scalar(S_PRES)] = 2.5;
        x += region == 1 ? 2.0 * dx : dx;
        if (x > 1.0) {
            y += dx;
            region = (y > 0.25 && y < 0.75) ? 2 : 1;
            if (region == 1) {
                if (odd) { odd = false; x = dx * 1.5; }
                else { odd = true; x = dx * 0.5; }
            } else x = dx * 0.5;
        }
    }
    finish_ideal_gas(q, gamma);
}

// N x N lattice on [-0.5, 0.5]^2, x fastest, running sums (gresho_chan_vortex.cpp:39-71, pairing_instability.cpp:22-50)
template <class F> void square_lattice(Particles & q, int N, F && fill)
{
    const int num = N * N;
    const double dx = 1.0 / N;
    q.resize(2, num);
    double x = -0.5 + dx * 0.5, y = -0.5 + dx * 0.5;
    for (int i = 0; i < num; ++i) {
        fill(q.rec_d(i), x, y);
        x += dx;
        if (x > 0.5) { x = -0.5 + dx * 0.5; y += dx; }
    }
}

void make_gresho(Particles & q, int N, double gamma)
{
    const double mass = 1.0 / ((double)N * N);
    square_lattice(q, N, [&](double * r, double x, double y) {
        const double rad = std::sqrt(x * x + y * y);
        double vel, pres;
        if (rad < 0.2) { vel = 5.0 * rad; pres = 5.0 + 12.5 * rad * rad; }
        else if (rad < 0.4) { vel = 2.0 - 5.0 * rad; pres = 9.0 + 12.5 * rad * rad - 20.0 * rad + 4.0 * std::log(5.0 * rad); }
        else { vel = 0.0; pres = 3.0 + 4.0 * std::log(2.0); }
        r[0] = x; r[1] = y;
        r[q.o_vel()] = (-y / rad) * vel;
        r[q.o_vel() + 1] = (x / rad) * vel;
        r[q.o_scalar(S_DENS)] = 1.0; r[q.o_scalar(S_PRES)] = pres; r[q.o_scalar(S_MASS)] = mass;
    });
    finish_ideal_gas(q, gamma);
}

void make_pairing(Particles & q, int N, double gamma)
{
    const double dx = 1.0 / N, mass = 1.0 / ((double)N * N);
    std::mt19937 engine(1);
    std::uniform_real_distribution<double> dist(-dx * 0.05, dx * 0.05);
    square_lattice(q, N, [&](double * r, double x, double y) {
        r[0] = x + dist(engine);
        r[1] = y + dist(engine);
        r[q.o_scalar(S_DENS)] = 1.0; r[q.o_scalar(S_PRES)] = 1.0; r[q.o_scalar(S_MASS)] = mass;
    });
    finish_ideal_gas(q, gamma);
}

// src/sample/hydrostatic.cpp:14-76: dense square (rho = 4) inside a thin ambient (rho = 1), equal masses
void make_hydrostatic(Particles & q, int N, double gamma)
{
    const double dx1 = 0.5 / N, dx2 = dx1 * 2.0, mass = 1.0 / ((double)N * N);
    std::vector<double> pts;                  // x, y, rho
    double x = -0.25 + dx1 * 0.5, y = -0.25 + dx1 * 0.5;
    while (y < 0.25) {
        pts.insert(pts.end(), {x, y, 4.0});
        x += dx1;
        if (x > 0.25) { x = -0.25 + dx1 * 0.5; y += dx1; }
    }
    x = -0.5 + dx2 * 0.5; y = -0.5 + dx2 * 0.5;
    while (y < 0.5) {
        pts.insert(pts.end(), {x, y, 1.0});
        do {
            x += dx2;
            if (x > 0.5) { x = -0.5 + dx2 * 0.5; y += dx2; }
        } while (x > -0.25 && x < 0.25 && y > -0.25 && y < 0.25);
    }
    const size_t n = pts.size() / 3;
    q.resize(2, n);
    for (size_t i = 0; i < n; ++i) {
        double * r = q.rec_d(i);
        r[0] = pts[3 * i]; r[1] = pts[3 * i + 1];
        r[q.o_scalar(S_MASS)] = mass; r[q.o_scalar(S_DENS)] = pts[3 * i + 2]; r[q.o_scalar(S_PRES)] = 2.5;
    }
    finish_ideal_gas(q, gamma);
}

// src/sample/evrard.cpp:19-63: N^3 lattice on [-1,1]^3 clipped to r <= 1, stretched r -> r^1.5 (rho ~ 1/r),
// i outermost, k innermost.  Two passes over the i-slabs with all host threads: count, then fill.
void make_evrard(Particles & q, int N, double gamma, double G)
{
    const double dx = 2.0 / N;
    std::vector<size_t> first(N + 1, 0);
    auto inside = [&](int i, int j, int k, double (&r)[3], double & r0) {
        r[0] = (i + 0.5) * dx - 1.0; r[1] = (j + 0.5) * dx - 1.0; r[2] = (k + 0.5) * dx - 1.0;
        r0 = std::sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
        return !(r0 > 1.0);
    };
#pragma omp parallel for schedule(dynamic, 1)
    for (int i = 0; i < N; ++i) {
        size_t c = 0;
        double r[3], r0;
        for (int j = 0; j < N; ++j) for (int k = 0; k < N; ++k) c += inside(i, j, k, r, r0) ? 1 : 0;
        first[i + 1] = c;
    }
    for (int i = 0; i < N; ++i) first[i + 1] += first[i];
    const size_t n = first[N];
    q.resize(3, n);
    const double mass = 1.0 / (double)n, u = 0.05 * G;
#pragma omp parallel for schedule(dynamic, 1)
    for (int i = 0; i < N; ++i) {
        size_t at = first[i];
        for (int j = 0; j < N; ++j) for (int k = 0; k < N; ++k) {
            double r[3], r0;
            if (!inside(i, j, k, r, r0)) continue;
            if (r0 > 0.0) {
                const double s = std::pow(r0, 1.5) / r0;
                r[0] *= s; r[1] *= s; r[2] *= s;
            }
            double * rec = q.rec_d(at);
            rec[0] = r[0]; rec[1] = r[1]; rec[2] = r[2];
            const double dens = 1.0 / (2.0 * M_PI * std::sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]));
            rec[q.o_scalar(S_MASS)] = mass; rec[q.o_scalar(S_DENS)] = dens; rec[q.o_scalar(S_ENE)] = u;
            rec[q.o_scalar(S_PRES)] = (gamma - 1.0) * dens * u;
            q.rec_i(at)[0] = (int)at;
            ++at;
        }
    }
}

void make_initial_condition(const Run & r, Particles & q)
{
    const std::string name = r.sample->name;
    const int N = r.n_side;
    if (name == "shock_tube") make_shock_tube(q, N, r.p.gamma);
    else if (name == "khi") make_khi(q, N, r.p.gamma);
    else if (name == "gresho_chan_vortex") make_gresho(q, N, r.p.gamma);
    else if (name == "pairing_instability") make_pairing(q, N, r.p.gamma);
    else if (name == "hydrostatic") make_hydrostatic(q, N, r.p.gamma);
    else if (name == "evrard") make_evrard(q, N, r.p.gamma, r.p.G);
    else throw std::runtime_error("unknown sample type.");
}

// ---------------------------------------------------------------------------------------------
// logging and output (src/logger.cpp, src/output.cpp)
// ---------------------------------------------------------------------------------------------
struct Log {
    std::ofstream file;
    void open(const std::string & dir)
    {
        struct stat st;
        bool made = false;
        if (stat(dir.c_str(), &st)) {
            // the reference creates one level only; create the parents too
            for (size_t i = 1; i <= dir.size(); ++i)
                if (i == dir.size() || dir[i] == '/') mkdir(dir.substr(0, i).c_str(), 0775);
            if (stat(dir.c_str(), &st)) throw std::runtime_error("cannot open directory");
            made = true;
        }
        std::time_t now = std::time(nullptr);
        std::tm * t = std::localtime(&now);
        char name[64];
        std::snprintf(name, sizeof(name), "/%04d%02d%02d%02d%02d%02d.log", t->tm_year + 1900, t->tm_mon + 1, t->tm_mday,
                      t->tm_hour, t->tm_min, t->tm_sec);
        file.open(dir + name);
        if (made) file << "mkdir " << dir << std::endl;
    }
    void line(const std::string & s, bool console = true)
    {
        if (console) std::cout << s << std::endl;
        if (file.is_open()) file << s << std::endl;
    }
};

#define CKB(ctx, call) do { if ((call) != 0) throw std::runtime_error(std::string("libsphb: ") + sphb_last_error(ctx)); } while (0)

struct Output {
    std::string dir;
    std::ofstream energy;
    int count = 0;
    bool binary = false;
    void open(const std::string & d)
    {
        dir = d;
        energy.open(dir + "/energy.dat");
        energy << "# time kinetic thermal potential total\n";
    }
    // src/output.cpp:14-64: pos vel acc mass dens pres ene sml id neighbor alpha gradh, default ostream precision
    void particles(sphb_ctx * c, Particles & q, double time, Log & log)
    {
        CKB(c, sphb_download_aos(c, q.bytes.data(), (int)q.size(), q.rec, SPHB_F_ALL));
        char name[32];
        std::snprintf(name, sizeof(name), "/%05d.dat", count);
        const std::string file = dir + name;
        std::ofstream out(file);
        out << "# " << time << std::endl;
        const int D = q.dim;
        for (size_t i = 0; i < q.size(); ++i) {
            const double * r = q.rec_d(i);
            for (int d = 0; d < D; ++d) out << r[q.o_pos() + d] << ' ';
            for (int d = 0; d < D; ++d) out << r[q.o_vel() + d] << ' ';
            for (int d = 0; d < D; ++d) out << r[q.o_acc() + d] << ' ';
            out << r[q.o_scalar(S_MASS)] << ' ' << r[q.o_scalar(S_DENS)] << ' ' << r[q.o_scalar(S_PRES)] << ' '
                << r[q.o_scalar(S_ENE)] << ' ' << r[q.o_scalar(S_SML)] << ' ' << q.rec_i(i)[0] << ' ' << q.rec_i(i)[1] << ' '
                << r[q.o_scalar(S_ALPHA)] << ' ' << r[q.o_scalar(S_GRADH)] << ' ' << '\n';
        }
        log.line("write " + file);
        if (binary) {
            // full-precision snapshot: header + the SPHParticle array as downloaded
            std::snprintf(name, sizeof(name), "/%05d.bin", count);
            std::ofstream bin(dir + name, std::ios::binary);
            const char magic[4] = {'S', 'P', 'H', 'B'};
            const int32_t dim = q.dim;
            const int64_t n = (int64_t)q.size(), rec = (int64_t)q.rec;
            bin.write(magic, 4);
            bin.write(reinterpret_cast<const char *>(&dim), 4);
            bin.write(reinterpret_cast<const char *>(&n), 8);
            bin.write(reinterpret_cast<const char *>(&rec), 8);
            bin.write(reinterpret_cast<const char *>(&time), 8);
            bin.write(reinterpret_cast<const char *>(q.bytes.data()), (std::streamsize)q.bytes.size());
            log.line("write " + dir + name);
        }
        ++count;
    }
    // src/output.cpp:66-90; the sums run on the device (sphb_energy)
    void energies(sphb_ctx * c, double time)
    {
        double e[3];
        CKB(c, sphb_energy(c, e));
        energy << time << " " << e[0] << " " << e[1] << " " << e[2] << " " << (e[0] + e[1] + e[2]) << std::endl;
    }
};

std::string fmt(const char * f, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, f);
    std::vsnprintf(buf, sizeof(buf), f, ap);
    va_end(ap);
    return buf;
}

} // namespace

int main(int argc, char ** argv)
{
    std::cout << "--------------SPH simulation-------------\n\n";
    std::string target, dump_ic, dump_params;
    std::vector<std::pair<std::string, std::string>> overrides;
    bool snapshots = true, binary = false;
    long max_steps = -1;
    int threads = 0, device = 0;
    for (int a = 1; a < argc; ++a) {
        const std::string s = argv[a];
        if (s == "--set" && a + 1 < argc) {
            const std::string kv = argv[++a];
            const size_t eq = kv.find('=');
            if (eq == std::string::npos) { std::cerr << "--set needs key=value" << std::endl; return EXIT_FAILURE; }
            overrides.emplace_back(kv.substr(0, eq), kv.substr(eq + 1));
        } else if (s == "--no-snapshots") snapshots = false;
        else if (s == "--binary-snapshots") binary = true;
        else if (s == "--steps" && a + 1 < argc) max_steps = std::atol(argv[++a]);
        else if (s == "--device" && a + 1 < argc) device = std::atoi(argv[++a]);
        else if (s == "--dump-ic" && a + 1 < argc) dump_ic = argv[++a];
        else if (s == "--dump-params" && a + 1 < argc) dump_params = argv[++a];
        else if (target.empty()) target = s;
        else threads = std::atoi(s.c_str());
    }
    if (target.empty()) {
        std::cerr << "how to use\n" << std::endl;
        std::cerr << "sph_gpu <paramter.json | sample name> [threads] [--set key=value]... [--no-snapshots] [--binary-snapshots] [--steps n] [--device d]" << std::endl;
        return EXIT_FAILURE;
    }
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#endif
    (void)threads;

    sphb_ctx * ctx = nullptr;
    Log log;
    try {
        const Run run = resolve(target, overrides);
        if (!dump_params.empty()) {
            // the resolved SPHParameters (+ time block, N, output directory), one "key value" per line, doubles with 17 digits:
            // compared with the reference's own Solver::read_parameterfile by tests/test_host_cpu.py
            std::ofstream f(dump_params);
            const sphb_params & P = run.p;
            f << "outputDirectory " << run.output_dir << "\n";
            f << fmt("startTime %.17g\nendTime %.17g\noutputTime %.17g\nenergyTime %.17g\n", run.t_start, run.t_end, run.t_output, run.t_energy);
            f << fmt("N %d\nsph_type %d\nkernel %d\ncfl_sound %.17g\ncfl_force %.17g\nav_alpha %.17g\n", run.n_side, P.sph_type, P.kernel, P.cfl_sound, P.cfl_force, P.av_alpha);
            f << fmt("use_balsara_switch %d\nuse_time_dependent_av %d\nalpha_max %.17g\nalpha_min %.17g\nepsilon_av %.17g\n", P.use_balsara_switch,
                     P.use_time_dependent_av, P.alpha_max, P.alpha_min, P.epsilon_av);
            f << fmt("use_ac %d\nalpha_ac %.17g\nmax_tree_level %d\nleaf_particle_num %d\nneighbor_number %d\niterative_sml %d\ngamma %.17g\n", P.use_ac, P.alpha_ac,
                     P.max_tree_level, P.leaf_particle_num, P.neighbor_number, P.iterative_sml, P.gamma);
            f << fmt("periodic %d\nuse_gravity %d\nG %.17g\ntheta %.17g\ngsph_2nd_order %d\n", P.periodic, P.use_gravity, P.G, P.theta, P.gsph_2nd_order);
            for (int d = 0; d < run.sample->dim; ++d) f << fmt("range_max%d %.17g\nrange_min%d %.17g\n", d, P.range_max[d], d, P.range_min[d]);
            if (dump_ic.empty()) return 0;
        }
        Particles q;
        make_initial_condition(run, q);
        if (!dump_ic.empty()) {
            std::ofstream f(dump_ic, std::ios::binary);
            f.write(reinterpret_cast<const char *>(q.bytes.data()), (std::streamsize)q.bytes.size());
            std::cout << "wrote " << q.size() << " particles (" << q.rec << " bytes each, DIM = " << q.dim << ") to " << dump_ic << std::endl;
            return 0;
        }
        log.open(run.output_dir);
        log.line("parameters");
        log.line("output directory     = " + run.output_dir);
        log.line("time");
        log.line(fmt("* start time         = %g", run.t_start));
        log.line(fmt("* end time           = %g", run.t_end));
        log.line(fmt("* output time        = %g", run.t_output));
        log.line(fmt("* enerty output time = %g", run.t_energy));
        log.line(std::string("SPH type: ") + (run.p.sph_type == SPHB_SSPH ? "Standard SPH" : run.p.sph_type == SPHB_DISPH ? "Density Independent SPH"
                 : run.p.gsph_2nd_order ? "Godunov SPH (2nd order)" : "Godunov SPH (1st order)"));
        log.line(fmt("Sample: %s, DIM = %d, N = %d, particles = %zu", run.sample->name, run.sample->dim, run.n_side, q.size()));
        log.line("device path: libsphb.so (sm_100a), state resident in HBM\n");

        if (sphb_create(&run.p, run.sample->dim, device, &ctx) != 0) throw std::runtime_error(std::string("libsphb: ") + sphb_last_error(nullptr));
        CKB(ctx, sphb_upload_aos(ctx, q.bytes.data(), (int)q.size(), q.rec, SPHB_F_ALL));
        CKB(ctx, sphb_initialize(ctx));                       // Solver::initialize, src/solver.cpp:353-414

        Output out;
        out.binary = binary;
        out.open(run.output_dir);
        double t = run.t_start, t_out = run.t_output, t_ene = run.t_energy;
        if (snapshots) out.particles(ctx, q, t, log);
        out.energies(ctx, t);

        const auto start = std::chrono::system_clock::now();
        auto t_cout_i = start;
        long loop = 0;
        while (t < run.t_end && (max_steps < 0 || loop < max_steps)) {
            double dt = 0.0;
            CKB(ctx, sphb_integrate(ctx, &dt));               // Solver::integrate, src/solver.cpp:417-429
            ++loop;
            t += dt;
            const auto now = std::chrono::system_clock::now();
            const std::string msg = fmt("loop: %ld, time: %g, dt: %g, num: %zu", loop, t, dt, q.size());
            if (std::chrono::duration_cast<std::chrono::seconds>(now - t_cout_i).count() >= 1) { log.line(msg); t_cout_i = now; }
            else log.line(msg, false);
            if (t > t_out) { if (snapshots) out.particles(ctx, q, t, log); t_out += run.t_output; }
            if (t > t_ene) { out.energies(ctx, t); t_ene += run.t_energy; }
        }
        CKB(ctx, sphb_synchronize(ctx));
        const auto end = std::chrono::system_clock::now();
        const double ms = (double)std::chrono::duration_cast<std::chrono::milliseconds>(end - start).count();
        log.line("\ncalculation is finished");
        log.line(fmt("calclation time: %g ms", ms));
        if (loop > 0 && ms > 0) log.line(fmt("particle-steps/s: %.4g (%ld steps)", (double)q.size() * loop / (ms * 1e-3), loop));
        if (sphb_nonconverged(ctx)) log.line(fmt("Newton iterations that fell back to the guess: %llu", (unsigned long long)sphb_nonconverged(ctx)));
        sphb_destroy(ctx);
    } catch (const std::exception & e) {
        // the reference's exception_handler: log, exit(EXIT_FAILURE) (include/exception.hpp:58-88)
        std::cerr << "error: " << e.what() << std::endl;
        if (log.file.is_open()) log.file << "error: " << e.what() << std::endl;
        if (ctx) sphb_destroy(ctx);
        return EXIT_FAILURE;
    }
    return 0;
}
