// cv_main.cpp -- host driver with the reference's CLI and file formats, running the constrained decode on
// the GPU through the C ABI (include/cv_b200.h).
//
// Mirrors reference src/main.rs:25-136 (flags -i/-o/-n/-b/-p/-t/-s, input files `sequences`, `tags`,
// `test_tags`, `hmm.json`, output file `<out>/<prop>_<run>`), src/utils.rs:7-60 (text loaders),
// src/hmm/hmm.rs:242-264 (hmm.json, ndarray-serde layout, null -> -inf), constraints.rs:40-69
// (Constraints::from_tags) and viterbi_solver/utils.rs:62-177 (SuperSequence::from, recompute_constraints,
// reorder).  Training (-t/-s, hmm.rs:22-190) is outside the hot path: it is rejected.
// recompute_constraints for 0 < prop < 1 needs rand 0.8's StdRng(3019) stream: see rng_chacha12.h.
//
// Debug hook: CV_DUMP_INPUTS=<file> writes the assembled solver inputs (obs, start, comp, seq) as text and
// CV_DRY_RUN=1 stops before any GPU call (used by the CPU tests of the assembly logic).
#include <algorithm>
#include <array>
#include <charconv>
#include <cmath>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <limits>
#include <map>
#include <set>
#include <sstream>
#include <string>
#include <vector>

#include "../include/cv_b200.h"
#include "rng_chacha12.h"

static const double NEG_INF = -std::numeric_limits<double>::infinity();
[[noreturn]] static void die(const std::string &m) { std::cerr << "error: " << m << std::endl; std::exit(101); }  // Rust panic exit code

// ---- src/utils.rs:7-34 ----
static std::vector<std::vector<std::array<size_t, 2>>> load_sequences(const std::string &path)
{
    std::ifstream f(path);
    if (!f) die("cannot open " + path);
    std::vector<std::vector<std::array<size_t, 2>>> ret;
    std::vector<std::array<size_t, 2>> cur;
    bool have_last = false; size_t last = 0;
    std::string line;
    while (std::getline(f, line)) {
        std::vector<size_t> s;
        size_t p = 0;
        while (true) {                                   // split(" "): every token must parse as usize
            size_t q = line.find(' ', p);
            std::string tok = line.substr(p, q == std::string::npos ? std::string::npos : q - p);
            if (tok.empty() || tok.find_first_not_of("0123456789") != std::string::npos) die("parse error in " + path + ": '" + line + "'");
            s.push_back(std::stoull(tok));
            if (q == std::string::npos) break;
            p = q + 1;
        }
        if (have_last && s[0] != last) { ret.push_back(cur); cur.clear(); }
        have_last = true; last = s[0];
        std::array<size_t, 2> el{0, 0};
        for (size_t i = 1; i < s.size(); i++) {
            if (i - 1 >= 2) die("more than D=2 features (reference: array index panic)");
            el[i - 1] = s[i];
        }
        cur.push_back(el);
    }
    ret.push_back(cur);
    return ret;
}

// ---- src/utils.rs:36-60; -1 -> None (encoded as -1) ----
static std::vector<std::vector<long long>> load_tags(const std::string &path)
{
    std::ifstream f(path);
    if (!f) die("cannot open " + path);
    std::vector<std::vector<long long>> ret;
    std::vector<long long> cur;
    bool have_last = false; size_t last = 0;
    std::string line;
    while (std::getline(f, line)) {
        size_t q = line.find(' ');
        if (q == std::string::npos) die("parse error in " + path);
        size_t tid = std::stoull(line.substr(0, q));
        size_t q2 = line.find(' ', q + 1);
        std::string t1 = line.substr(q + 1, q2 == std::string::npos ? std::string::npos : q2 - q - 1);
        if (have_last && tid != last) { ret.push_back(cur); cur.clear(); }
        cur.push_back(t1 == "-1" ? -1 : (long long)std::stoull(t1));
        have_last = true; last = tid;
    }
    ret.push_back(cur);
    return ret;
}

// ---- minimal JSON reader for hmm.json (hmm.rs:242-264) ----
struct Json {
    const std::string &s; size_t p = 0;
    explicit Json(const std::string &str) : s(str) {}
    void ws() { while (p < s.size() && std::isspace((unsigned char)s[p])) p++; }
    bool eat(char c) { ws(); if (p < s.size() && s[p] == c) { p++; return true; } return false; }
    void expect(char c) { if (!eat(c)) die(std::string("hmm.json: expected '") + c + "' at " + std::to_string(p)); }
    std::string key() { ws(); expect('"'); size_t q = s.find('"', p); std::string k = s.substr(p, q - p); p = q + 1; expect(':'); return k; }
    double number_or_null()
    {
        ws();
        if (s.compare(p, 4, "null") == 0) { p += 4; return NEG_INF; }
        const char *b = s.c_str() + p; char *e = nullptr;
        double v = std::strtod(b, &e);
        if (e == b) die("hmm.json: bad number at " + std::to_string(p));
        p += (size_t)(e - b);
        return v;
    }
    // {"v":1,"dim":[...],"data":[numbers]}
    void ndarray(std::vector<size_t> &dim, std::vector<double> &data)
    {
        expect('{');
        do {
            std::string k = key();
            if (k == "v") number_or_null();
            else if (k == "dim") { expect('['); dim.clear(); if (!eat(']')) { do dim.push_back((size_t)number_or_null()); while (eat(',')); expect(']'); } }
            else if (k == "data") { expect('['); data.clear(); if (!eat(']')) { do data.push_back(number_or_null()); while (eat(',')); expect(']'); } }
            else die("hmm.json: unknown key " + k);
        } while (eat(','));
        expect('}');
    }
};

struct Hmm { int K = 0; std::vector<size_t> bdims; std::vector<double> a, b, pi; size_t M = 1; };

static Hmm hmm_from_json(const std::string &path)
{
    std::ifstream f(path);
    if (!f) die("cannot open " + path);
    std::stringstream ss; ss << f.rdbuf();
    std::string text = ss.str();
    Json j(text);
    Hmm h;
    j.expect('{');
    do {
        std::string k = j.key();
        std::vector<size_t> dim; std::vector<double> data;
        if (k == "a") { j.ndarray(dim, data); if (dim.size() != 2 || dim[0] != dim[1]) die("hmm.json: a must be [K,K]"); h.K = (int)dim[0]; h.a = data; }
        else if (k == "pi") { j.ndarray(dim, data); h.pi = data; }
        else if (k == "b") {      // Array1 of ArrayD: {"v":1,"dim":[K],"data":[ {ndarray}, ... ]}
            j.expect('{');
            do {
                std::string kk = j.key();
                if (kk == "v") j.number_or_null();
                else if (kk == "dim") { j.expect('['); j.number_or_null(); j.expect(']'); }
                else if (kk == "data") {
                    j.expect('[');
                    do { std::vector<size_t> d2; std::vector<double> blk; j.ndarray(d2, blk); h.bdims = d2; h.b.insert(h.b.end(), blk.begin(), blk.end()); } while (j.eat(','));
                    j.expect(']');
                } else die("hmm.json: unknown key " + kk);
            } while (j.eat(','));
            j.expect('}');
        } else die("hmm.json: unknown key " + k);
    } while (j.eat(','));
    j.expect('}');
    h.M = 1; for (size_t d : h.bdims) h.M *= d;
    if ((size_t)h.K * h.K != h.a.size() || (size_t)h.K != h.pi.size() || (size_t)h.K * h.M != h.b.size()) die("hmm.json: inconsistent shapes");
    return h;
}

// ---- Rust `{}` for f64: shortest round-trip digits, never exponent notation, -inf / inf / NaN ----
static std::string rust_f64(double v)
{
    if (std::isnan(v)) return "NaN";
    if (std::isinf(v)) return v < 0 ? "-inf" : "inf";
    char buf[512];
    auto r = std::to_chars(buf, buf + sizeof buf, v, std::chars_format::fixed);
    return std::string(buf, r.ptr);
}

struct Element { size_t seq, t; std::array<size_t, 2> value; int comp; bool active; };

int main(int argc, char **argv)
{
    std::string input, output = "."; int nstates = -1; std::vector<size_t> nobs; double prop = NAN; bool have_prop = false, train = false;
    std::string prop_text;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        auto need = [&](const char *n) -> std::string { if (i + 1 >= argc) die(std::string("missing value for ") + n); return argv[++i]; };
        if (a == "-i" || a == "--input") input = need("INPUT");
        else if (a == "-o" || a == "--output") output = need("OUTPUT");
        else if (a == "-n" || a == "--nstates") nstates = std::stoi(need("NSTATES"));
        else if (a == "-b" || a == "--nobs") { while (i + 1 < argc && argv[i + 1][0] != '-') nobs.push_back(std::stoull(argv[++i])); }
        else if (a == "-p" || a == "--prop") { prop_text = need("PROP"); prop = std::stod(prop_text); have_prop = true; }
        else if (a == "-t" || a == "--train") train = true;
        else if (a == "-s" || a == "--supervised") {}
        else die("unknown argument " + a);
    }
    if (input.empty() || nstates < 0 || !have_prop) die("required: -i INPUT -n NSTATES -p PROP (main.rs:31-58)");
    if (nobs.size() < 2) die("-b needs two values (main.rs:76,91 unwraps nobs[0], nobs[1])");
    if (train) die("-t/--train (HMM::new + MLE/EM, hmm.rs:22-190) is outside the GPU hot path; provide hmm.json");

    std::cout << "Loading data" << std::endl;                                   // main.rs:80
    auto sequences = load_sequences(input + "/sequences");                      // main.rs:82 (D = 2)
    auto tags = load_tags(input + "/tags"); (void)tags;                         // main.rs:84 (training only)
    auto control = load_tags(input + "/test_tags");                             // main.rs:86
    Hmm hmm = hmm_from_json(input + "/hmm.json");                               // main.rs:100-101
    if (hmm.K != nstates) std::cerr << "warning: -n " << nstates << " differs from hmm.json K=" << hmm.K << std::endl;
    if (hmm.bdims.size() != 2) die("hmm.json: b blocks must be 2-dimensional (D = 2, main.rs:82)");

    // ---- Constraints::from_tags (constraints.rs:40-69) ----
    std::vector<size_t> comp_value; std::vector<std::set<std::pair<size_t, size_t>>> components;
    for (size_t sid = 0; sid < control.size(); sid++)
        for (size_t t = 0; t < control[sid].size(); t++) {
            if (control[sid][t] < 0) continue;
            size_t tag = (size_t)control[sid][t];
            size_t id = std::find(comp_value.begin(), comp_value.end(), tag) - comp_value.begin();
            if (id == comp_value.size()) { comp_value.push_back(tag); components.emplace_back(); }
            components[id].insert({sid, t});
        }

    // ---- SuperSequence::from (viterbi_solver/utils.rs:62-103) ----
    std::vector<Element> el;
    std::vector<size_t> start(sequences.size()), sizes(sequences.size());
    for (size_t sid = 0; sid < sequences.size(); sid++) {
        start[sid] = el.size(); sizes[sid] = sequences[sid].size();
        for (size_t t = 0; t < sequences[sid].size(); t++) {
            int c = -1;
            for (size_t cid = 0; cid < components.size(); cid++) if (components[cid].count({sid, t})) { c = (int)cid; break; }
            el.push_back({sid, t, sequences[sid][t], c, c != -1});
        }
    }
    const size_t N = el.size();
    auto flat = [&](const Element &e) -> size_t {
        if (e.value[0] >= hmm.bdims[0] || e.value[1] >= hmm.bdims[1]) die("observation out of bounds (reference: ndarray index panic)");
        return e.value[0] * hmm.bdims[1] + e.value[1];
    };
    // ---- recompute_constraints(prop) + reorder (utils.rs:105-177); main.rs:107 and again main.rs:113-115 ----
    StdRng rng = StdRng::seed_from_u64(3019);                                   // utils.rs:101
    auto recompute = [&]() {
        for (auto &e : el) e.active = (e.comp != -1) && (rng.gen_f64() <= prop);            // short-circuit like Rust's &&
        // get_sequences_ordering: (last element active, avg #emittable states, seq id), stable ascending
        struct Key { int w; double avg; size_t id; };
        std::vector<Key> keys;
        for (size_t sid = 0; sid < start.size(); sid++) {
            double possible = 0.0; bool cons = false;
            for (size_t i = start[sid]; i < start[sid] + sizes[sid]; i++) {
                cons = el[i].active;
                size_t o = flat(el[i]);
                for (int s = 0; s < hmm.K; s++) if (hmm.b[(size_t)s * hmm.M + o] > NEG_INF) possible += 1.0;
            }
            keys.push_back({cons ? 1 : 0, possible / (double)sizes[sid], sid});
        }
        std::stable_sort(keys.begin(), keys.end(), [](const Key &x, const Key &y) {
            if (x.w != y.w) return x.w < y.w;
            if (x.avg != y.avg) return x.avg < y.avg;     // NaN (empty sequence) would panic in the reference
            return x.id < y.id;
        });
        std::vector<Element> ne; ne.reserve(N);
        std::vector<size_t> nstart = start;
        for (auto &k : keys) { nstart[k.id] = ne.size(); for (size_t i = start[k.id]; i < start[k.id] + sizes[k.id]; i++) ne.push_back(el[i]); }
        el.swap(ne); start.swap(nstart);
    };
    recompute();                                                                // main.rs:107
    if (prop != 0.0 && prop != 1.0) recompute();                                // main.rs:113-115 (run 0)
    std::set<int> act; for (auto &e : el) if (e.active) act.insert(e.comp);
    const int ncomp = (int)act.size();                                          // number_constraints()

    std::vector<uint32_t> obs(N); std::vector<uint8_t> st(N); std::vector<int32_t> comp(N);
    for (size_t i = 0; i < N; i++) { obs[i] = (uint32_t)flat(el[i]); st[i] = el[i].t == 0; comp[i] = el[i].active ? el[i].comp : -1; }
    if (const char *dump = std::getenv("CV_DUMP_INPUTS")) {
        std::ofstream d(dump);
        d << N << " " << ncomp << "\n";
        for (size_t i = 0; i < N; i++) d << el[i].seq << " " << obs[i] << " " << (int)st[i] << " " << comp[i] << "\n";
    }
    if (std::getenv("CV_DRY_RUN")) return 0;

    // ---- solver (main.rs:120-126) ----
    cv_hmm *h = nullptr;
    std::vector<uint64_t> bd(hmm.bdims.begin(), hmm.bdims.end());
    if (cv_hmm_create(hmm.K, 2, bd.data(), hmm.a.data(), hmm.b.data(), hmm.pi.data(), -1, &h)) die(cv_last_error());
    std::cout << "[cp EXP " << rust_f64(prop) << "] Run 1/1" << std::endl;     // main.rs:123
    std::vector<uint64_t> sol(N); double obj = 0; uint64_t explored = 0, steps = 0;
    auto t0 = std::chrono::steady_clock::now();
    uint64_t max_nodes = 0; if (const char *mn = std::getenv("CV_MAX_NODES")) max_nodes = std::strtoull(mn, nullptr, 10);
    if (cv_cp_solve(h, obs.data(), st.data(), comp.data(), (int64_t)N, ncomp, max_nodes, sol.data(), &obj, &explored, &steps)) die(cv_last_error());
    auto ms = std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::steady_clock::now() - t0).count();

    // ---- output (main.rs:111,129-133) ----
    std::ofstream out(output + "/" + rust_f64(prop) + "_0");
    if (!out) die("cannot create output file");
    out << rust_f64(obj) << " " << explored << "\n" << ms << "\n";
    for (size_t i = 0; i < N; i++) out << el[i].seq << " " << sol[i] << "\n";
    cv_hmm_destroy(h);
    return 0;
}
