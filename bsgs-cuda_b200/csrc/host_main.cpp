// host_main.cpp -- the host program above the C ABI, in C++ (the reference host is compiled code:
// PureBasic, 1_9_7File.pb; its toolchain is not available here, see INTEGRATION.md for the binding a
// PureBasic maintainer would write).  Same roles, same flags, same files, same console lines as the
// reference's main program, driving libbsgs_b200.so only through include/bsgs_b200.h:
//
//   getprogparam   1_9_7File.pb:875-1042   -> parse_args / validate
//   getCudaDevices + Tune     778-860, 324-431 -> list_devices
//   Save_HTpacked / LOAD_HTGPUpacked / Save_Load_Giants  3645-3814, 1905-2058 -> load_or_build_tables
//   main set-up    4689-4765, 5042-5064    -> Search
//   GetJob         2077-2092               -> bsgs_run (P_j = P0 + j*D inside the library)
//   cuda()         2360-2535 (one thread per GPU, 4776) -> worker
//   checkerThread  3933-4296               -> check_record (records arrive verified, with the payload j)
//   saveCurentCNT  3897-3931               -> save_checkpoint
//   win.txt        5151-5157
//
// bsgs-cuda_b200/cli.py is the same program in Python; the two are tested against each other.
#include "../../include/bsgs_b200.h"
#include "host_ec.h"
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace {

// ------------------------------------------------------------------------------------------ 256-bit scalars mod n
struct U256 { uint64_t v[4] = {0, 0, 0, 0}; };
const U256 kN = {{0xBFD25E8CD0364141ULL, 0xBAAEDCE6AF48A03BULL, 0xFFFFFFFFFFFFFFFEULL, 0xFFFFFFFFFFFFFFFFULL}};
typedef unsigned __int128 u128;

int cmp(const U256& a, const U256& b) { for (int i = 3; i >= 0; i--) if (a.v[i] != b.v[i]) return a.v[i] < b.v[i] ? -1 : 1; return 0; }
bool is_zero(const U256& a) { return (a.v[0] | a.v[1] | a.v[2] | a.v[3]) == 0; }
uint64_t add_raw(U256& r, const U256& a, const U256& b) { u128 c = 0; for (int i = 0; i < 4; i++) { c += (u128)a.v[i] + b.v[i]; r.v[i] = (uint64_t)c; c >>= 64; } return (uint64_t)c; }
uint64_t sub_raw(U256& r, const U256& a, const U256& b) { uint64_t br = 0; for (int i = 0; i < 4; i++) { u128 d = (u128)a.v[i] - b.v[i] - br; r.v[i] = (uint64_t)d; br = (uint64_t)(d >> 64) & 1; } return br; }
U256 from_u64(uint64_t x) { U256 r; r.v[0] = x; return r; }
U256 mod_n(U256 a) { while (cmp(a, kN) >= 0) sub_raw(a, a, kN); return a; }
U256 addmod(const U256& a, const U256& b) { U256 r; uint64_t c = add_raw(r, a, b); if (c || cmp(r, kN) >= 0) sub_raw(r, r, kN); return r; }      // a, b < n
U256 submod(const U256& a, const U256& b) { U256 r; if (sub_raw(r, a, b)) add_raw(r, r, kN); return r; }
U256 negmod(const U256& a) { return is_zero(a) ? a : submod(U256(), a); }
// a * k for results below 2^256 (window counters: j * 4Nw < 2^130)
U256 mul_small(const U256& a, uint64_t k) { U256 r; u128 c = 0; for (int i = 0; i < 4; i++) { c += (u128)a.v[i] * k; r.v[i] = (uint64_t)c; c >>= 64; } return r; }

bool parse_hex(const std::string& s, U256& out) {
    if (s.empty() || s.size() > 64) return false;
    out = U256();
    for (char ch : s) {
        int d = ch >= '0' && ch <= '9' ? ch - '0' : ch >= 'a' && ch <= 'f' ? ch - 'a' + 10 : ch >= 'A' && ch <= 'F' ? ch - 'A' + 10 : -1;
        if (d < 0) return false;
        for (int i = 3; i > 0; i--) out.v[i] = (out.v[i] << 4) | (out.v[i - 1] >> 60);
        out.v[0] = (out.v[0] << 4) | (uint64_t)d;
    }
    return true;
}
std::string hex64(const uint64_t v[4]) { char b[65]; snprintf(b, sizeof b, "%016llx%016llx%016llx%016llx", (unsigned long long)v[3], (unsigned long long)v[2], (unsigned long long)v[1], (unsigned long long)v[0]); return b; }
std::string hex_short(const U256& a) { std::string h = hex64(a.v); size_t p = h.find_first_not_of('0'); return p == std::string::npos ? "0" : h.substr(p); }

// ------------------------------------------------------------------------------------------ curve helpers on top of host_ec.h
using hec::Fe; using hec::Pt;
Fe fe_of(const U256& a) { Fe r; memcpy(r.v, a.v, 32); return r; }
// k*a for a 256-bit k >= 1, LSB-first double-and-add like m_PTMULX64 (Curve64.pb:2586-2619)
Pt pt_mul(const Pt& a, const U256& k) {
    Pt acc = a, res = a; bool have = false;
    int top = 255; while (top > 0 && !((k.v[top / 64] >> (top % 64)) & 1)) top--;
    for (int i = 0; i <= top; i++) {
        if ((k.v[i / 64] >> (i % 64)) & 1) { if (have) hec::pt_add(res, res, acc); else { res = acc; have = true; } }
        if (i < top) hec::pt_add(acc, acc, acc);
    }
    return res;
}
Pt pt_sub(const Pt& a, const Pt& b) { Pt nb, r; hec::pt_neg(nb, b); hec::pt_add(r, a, nb); return r; }
// y from x (m_YfromX64, Curve64.pb:2656-2682): y = (x^3 + 7)^((p+1)/4), parity fixed by the prefix
bool y_from_x(const Fe& x, int odd, Fe& y) {
    Fe t, rhs, seven = {{7, 0, 0, 0}};
    hec::mul(t, x, x); hec::mul(t, t, x); hec::add(rhs, t, seven);
    const uint64_t e[4] = {0xFFFFFFFFBFFFFF0CULL, ~0ULL, ~0ULL, 0x3FFFFFFFFFFFFFFFULL};      // (p+1)/4
    Fe acc = {{1, 0, 0, 0}};
    for (int i = 255; i >= 0; i--) { hec::mul(acc, acc, acc); if ((e[i / 64] >> (i % 64)) & 1) hec::mul(acc, acc, rhs); }
    hec::mul(t, acc, acc);
    if (!hec::eq(t, rhs)) return false;
    if ((int)(acc.v[0] & 1) != odd) hec::neg(acc, acc);
    y = acc; return true;
}
std::string lower(std::string s) { for (auto& c : s) c = (char)tolower((unsigned char)c); return s; }
std::string cut_hex(std::string s) {                                     // cutHex, 1_9_7File.pb:864-873
    while (!s.empty() && isspace((unsigned char)s.back())) s.pop_back();
    size_t b = 0; while (b < s.size() && isspace((unsigned char)s[b])) b++;
    s = lower(s.substr(b));
    return s.rfind("0x", 0) == 0 ? s.substr(2) : s;
}
// 128 hex | '04'+128 | '02'/'03'+64 (1_9_7File.pb:5006-5018)
bool parse_pubkey(std::string s, Pt& out) {
    s = cut_hex(s);
    if (s.size() == 130 && s.rfind("04", 0) == 0) s = s.substr(2);
    U256 x, y;
    if (s.size() == 128) { if (!parse_hex(s.substr(0, 64), x) || !parse_hex(s.substr(64), y)) return false; out.x = fe_of(x); out.y = fe_of(y); return true; }
    if (s.size() == 66 && (s.rfind("02", 0) == 0 || s.rfind("03", 0) == 0)) {
        if (!parse_hex(s.substr(2), x)) return false;
        out.x = fe_of(x);
        return y_from_x(out.x, s[1] == '3', out.y);
    }
    return false;
}
std::string compress(const Pt& p) { return std::string((p.y.v[0] & 1) ? "03" : "02") + hex64(p.x.v); }

// ------------------------------------------------------------------------------------------ SHA-1 (settings fingerprint, 1_9_7File.pb:4634-4636)
std::string sha1_hex(const std::string& msg) {
    uint32_t h[5] = {0x67452301u, 0xEFCDAB89u, 0x98BADCFEu, 0x10325476u, 0xC3D2E1F0u};
    std::string m = msg; uint64_t bits = (uint64_t)msg.size() * 8;
    m.push_back((char)0x80); while (m.size() % 64 != 56) m.push_back(0);
    for (int i = 7; i >= 0; i--) m.push_back((char)(bits >> (8 * i)));
    auto rol = [](uint32_t x, int n) { return (x << n) | (x >> (32 - n)); };
    for (size_t off = 0; off < m.size(); off += 64) {
        uint32_t w[80];
        for (int i = 0; i < 16; i++) w[i] = ((uint32_t)(uint8_t)m[off + 4 * i] << 24) | ((uint32_t)(uint8_t)m[off + 4 * i + 1] << 16) | ((uint32_t)(uint8_t)m[off + 4 * i + 2] << 8) | (uint8_t)m[off + 4 * i + 3];
        for (int i = 16; i < 80; i++) w[i] = rol(w[i - 3] ^ w[i - 8] ^ w[i - 14] ^ w[i - 16], 1);
        uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4];
        for (int i = 0; i < 80; i++) {
            uint32_t f, k;
            if (i < 20) { f = (b & c) | (~b & d); k = 0x5A827999u; } else if (i < 40) { f = b ^ c ^ d; k = 0x6ED9EBA1u; }
            else if (i < 60) { f = (b & c) | (b & d) | (c & d); k = 0x8F1BBCDCu; } else { f = b ^ c ^ d; k = 0xCA62C1D6u; }
            uint32_t t = rol(a, 5) + f + e + k + w[i]; e = d; d = c; c = rol(b, 30); b = a; a = t;
        }
        h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e;
    }
    char out[41]; snprintf(out, sizeof out, "%08x%08x%08x%08x%08x", h[0], h[1], h[2], h[3], h[4]);
    return out;
}

// ------------------------------------------------------------------------------------------ options (getprogparam, 1_9_7File.pb:875-1042)
struct Options {
    uint32_t t = 256, b = 132, p = 400, htsz = 25; uint64_t w = 1ull << 25; int wt = 180, sf = 1;    // 1_9_7File.pb:181-186, 168, 152
    std::string pb = "036d05521c67b9cc1c0ef906b42215c7120c7302c34d9316a2726199bedac50936", d = "0", infile, wl;
    std::string pk_str = "01", pke_str = "1ffffffffffffffff";
    U256 pk = from_u64(1), pke; bool nosave = false, fingerprint_only = false, onlygen = false;
    std::vector<int> devices;
};
struct UsageError { std::string msg; };
void print_try(unsigned long long free_bytes, int sm, int major, int minor);      // Tune listing, below

uint64_t parse_w(const std::string& arg) {                                // 1_9_7File.pb:1013-1021: <= 32 means Int(2^x)
    double v = atof(arg.c_str());
    return v <= 32 ? (uint64_t)pow(2.0, v) : strtoull(arg.c_str(), nullptr, 10);
}
std::string format_status(const std::string& cnt_hex, int running, const std::vector<unsigned long long>& rates, uint64_t w, double job_s, double total_s);
Options parse_args(const std::vector<std::string>& argv) {
    Options o; parse_hex("1ffffffffffffffff", o.pke);
    for (size_t i = 0; i < argv.size(); i++) {
        std::string a = lower(argv[i]);
        bool has = i + 1 < argv.size() && !(argv[i + 1].size() > 1 && argv[i + 1][0] == '-' && !isdigit((unsigned char)argv[i + 1][1]));
        const std::string val = has ? argv[i + 1] : "";
        if (a == "-h") throw UsageError{"usage: bsgs_b200 -t T -b B -p P -w W -htsz H -pb PUB -pk HEX -pke HEX [-d 0,1] [-infile F] [-wl F] [-wt S] [-nosave]"};
        if (a == "-nosave") { o.nosave = true; continue; }
        if (a == "-onlygen") { o.onlygen = true; continue; }                 // the role of onlygen1_9_6File.pb: write the table files and stop
        if (a == "-fingerprint") { o.fingerprint_only = true; continue; }    // extension used by the tests: print the settings hash and exit
        if (a == "-tune") {                                                  // extension used by the tests: -tune free_bytes,sm,major,minor prints the Try lines
            unsigned long long mem = 0; int sm = 0, mj = 0, mn = 0;
            if (!has || sscanf(val.c_str(), "%llu,%d,%d,%d", &mem, &sm, &mj, &mn) != 4) throw UsageError{"-tune free_bytes,sm,major,minor"};
            print_try(mem, sm, mj, mn); exit(0);
        }
        if (a == "-statusline") {                                            // extension used by the tests: -statusline cnthex,running,w,job_s,total_s,rate0[,rate1...]
            std::vector<std::string> f; std::string v = val + ","; size_t q;
            while ((q = v.find(',')) != std::string::npos) { f.push_back(v.substr(0, q)); v.erase(0, q + 1); }
            if (f.size() < 6) throw UsageError{"-statusline cnthex,running,w,job_s,total_s,rate0[,rate1...]"};
            std::vector<unsigned long long> rates; for (size_t k = 5; k < f.size(); k++) rates.push_back(strtoull(f[k].c_str(), nullptr, 10));
            puts(format_status(f[0], atoi(f[1].c_str()), rates, strtoull(f[2].c_str(), nullptr, 10), atof(f[3].c_str()), atof(f[4].c_str())).c_str()); exit(0);
        }
        if (a == "-t") { if (has) { o.t = (uint32_t)atoi(val.c_str()); i++; } }
        else if (a == "-b") { if (has) { o.b = (uint32_t)atoi(val.c_str()); i++; } }
        else if (a == "-p") { if (has) { o.p = (uint32_t)atoi(val.c_str()); i++; } }
        else if (a == "-d") { if (has) { o.d = val; i++; } }
        else if (a == "-htsz") { if (has) { o.htsz = (uint32_t)atoi(val.c_str()); i++; } }
        else if (a == "-infile") { if (has) { o.infile = val; i++; } }
        else if (a == "-wl") { if (has) { o.wl = val; i++; } }
        else if (a == "-sf") { if (has) { o.sf = atoi(val.c_str()); i++; } }
        else if (a == "-pb") { if (has) { o.pb = cut_hex(val); i++; } }
        else if (a == "-pk") { if (has) { o.pk_str = cut_hex(val); if (!parse_hex(o.pk_str, o.pk)) throw UsageError{"bad -pk"}; i++; } }
        else if (a == "-pke") { if (has) { o.pke_str = cut_hex(val); if (!parse_hex(o.pke_str, o.pke)) throw UsageError{"bad -pke"}; i++; } }
        else if (a == "-w") { if (has) { o.w = parse_w(val); i++; } }
        else if (a == "-wt") { if (has) { o.wt = std::max(atoi(val.c_str()), 30); i++; } }                    // 1_9_7File.pb:935-937
        else throw UsageError{"Unknown parameter [" + argv[i] + "]"};                                         // 1_9_7File.pb:1033
    }
    size_t pos = 0; std::string d = o.d + ",";
    while ((pos = d.find(',')) != std::string::npos) { if (pos) o.devices.push_back(atoi(d.substr(0, pos).c_str())); d.erase(0, pos + 1); }
    return o;
}
std::vector<std::string> validate(const Options& o) {                    // 1_9_7File.pb:4412-4472, 4616-4618, 4904-4922
    if (o.w >= 3069485951ull) throw UsageError{"-w should be less than 3069485951"};
    if (o.htsz > 31) throw UsageError{"-htsz should be less or equal to 31"};
    if (o.p % 2) throw UsageError{"vaule -p must be a multiple of 2"};
    if (o.onlygen && (o.t % 2 || o.b % 2)) throw UsageError{"vaule -t and -b must be a multiple of 2"};      // onlygen1_9_6File.pb:3140-3145
    if (is_zero(o.pk)) throw UsageError{"Range begin should be more than 0"};
    if (cmp(o.pke, o.pk) <= 0) throw UsageError{"Range end should be more than range begin"};
    std::vector<std::string> warn;
    uint64_t lim = o.htsz == 27 ? 1331331443ull : o.htsz == 28 ? 1777178603ull : 0;
    if (lim && o.w > lim) warn.push_back("UNSAFE: with htsz " + std::to_string(o.htsz) + " value -w should be <= " + std::to_string(lim));
    if ((int)floor(log2((double)o.w)) - (int)o.htsz > 3) warn.push_back("Warning: hashtable is small for this -w, probes will be slow");
    return warn;
}
std::string fingerprint(const Options& o) {
    return sha1_hex(std::to_string(o.t) + std::to_string(o.b) + std::to_string(o.p) + std::to_string(o.w) + o.pk_str + o.pke_str + std::to_string(o.htsz));
}

// ------------------------------------------------------------------------------------------ Tune (1_9_7File.pb:324-431, core table 816-846)
int cores_for(int major, int minor, int mp) {
    if (major == 2) return mp * (minor == 1 ? 48 : 32);
    if (major == 3) return mp * 192;
    if (major == 5) return mp * 128;
    if (major == 6 || major == 7) return mp * 64;
    return major >= 8 ? mp * 128 : 0;                                      // reference stops at 8; 9 and 10 also have 128 FP32 lanes per SM
}
struct Try { long long t, b, p, k, htsz; double mb; bool ok; };
Try tune_reference(long long memsize, long long sm, long long cores) {
    long long blockN = sm * 2, treadN = 2 * cores / sm, pparamN = 140, htszN = 0, httot = 0;
    long long gsize = treadN * blockN * pparamN * 96 + 160 + 128;
    if (memsize - gsize <= 0) return Try{0, 0, 0, 0, 0, 0, false};
    long long k = (long long)floor(log((double)(memsize - gsize)) / log(2.0) + 0.5);     // float -> integer assignment rounds in PureBasic
    if (k > 31) k = 31;
    k++;
    do {
        k--;
        if (k == 31) htszN = 29; else if (k >= 28 && k <= 30) htszN = 27; else if (k == 27) htszN = 25;
        httot = (1ll << k) * 4 + (1ll << htszN) * 4;
    } while (!(gsize + httot < memsize || k < 25));
    long long frmem;
    do {
        frmem = (memsize - httot - 160 - 128) / 96 / blockN / treadN;
        if (frmem > 512) { if (treadN < 512) treadN += 32; else blockN += sm; }
    } while (!(frmem < 512 || blockN / sm >= 4));
    frmem = (memsize - httot - 160 - 128 - 96) / 96 / blockN / treadN;
    if (frmem & 1) frmem++;
    pparamN = frmem;
    gsize = treadN * blockN * pparamN * 96 + 160 + 128;
    if (gsize + httot > memsize) { pparamN -= 2; gsize = treadN * blockN * pparamN * 96 + 160 + 128; }
    return Try{treadN, blockN, pparamN, k, htszN, (double)(gsize + httot) / 1048576.0, true};
}
// this build's own suggestion (bsgs-cuda_b200/tune.py::tune_b200)
bool tune_b200(unsigned long long free_bytes, int sm, double& wexp, long long& p, int& htsz, double& mb) {
    const int windows = 16; const long long t = 512;
    const int steps[] = {63, 62, 61, 60, 59, 58, 57, 56, 54, 52, 50, 48};
    for (int w2 : steps) {
        double we = w2 / 2.0; unsigned long long w = (unsigned long long)pow(2.0, we);
        if (w >= (1ull << 32)) continue;
        int hs = std::min(31, (int)floor(we) - 2);
        unsigned long long H = 1ull << hs, tb = 4 * (H + 1) + 4 * w + 32 * H + 4 * w;
        if ((double)tb > 0.5 * (double)free_bytes) continue;
        unsigned long long room = (unsigned long long)((double)(free_bytes - tb) * 0.9);
        unsigned long long nmax = std::min<unsigned long long>(room / (64 + 2 * 32 * windows), (1ull << 28) - 1);      // two chains: two launches in flight
        long long pp = (long long)(nmax / (unsigned long long)(t * sm)) & ~1ll;
        if (pp < 2) continue;
        wexp = we; p = pp; htsz = hs; mb = (double)(tb + (unsigned long long)(t * sm * pp) * (64 + 2 * 32 * windows)) / 1048576.0;
        return true;
    }
    return false;
}
void logln(const std::string& s) { printf("%s\n", s.c_str()); fflush(stdout); }
void print_try(unsigned long long free_bytes, int sm, int major, int minor) {
    char buf[512];
    int cores = cores_for(major, minor, sm);
    snprintf(buf, sizeof buf, "Device have: MP:%d Cores+%d", sm, cores); logln(buf);
    if (!sm || !cores) return;
    Try r = tune_reference((long long)free_bytes, sm, cores);
    if (r.ok) { snprintf(buf, sizeof buf, "Try -t %lld -b %lld -p %lld -w %lld -htsz %lld [%.3f MB]   (reference Tune)", r.t, r.b, r.p, r.k, r.htsz, r.mb); logln(buf); }
    double we, mb; long long p; int hs;
    if (tune_b200(free_bytes, sm, we, p, hs, mb)) {
        if (we == floor(we)) snprintf(buf, sizeof buf, "Try -t 512 -b %d -p %lld -w %d -htsz %d [%.3f MB]   (this build)", sm, p, (int)we, hs, mb);
        else snprintf(buf, sizeof buf, "Try -t 512 -b %d -p %lld -w %.1f -htsz %d [%.3f MB]   (this build)", sm, p, we, hs, mb);
        logln(buf);
    }
}
void list_devices() {
    int n = 0;
    if (bsgs_device_count(&n) != BSGS_OK) { logln("device listing unavailable"); return; }
    for (int i = 0; i < n; i++) {
        bsgs_devinfo d;
        if (bsgs_device_query(i, &d) != BSGS_OK) continue;
        char buf[512];
        snprintf(buf, sizeof buf, "Cuda device:%s (%.3f/%lluMB)", d.name, (double)d.free_bytes / 1048576.0, (unsigned long long)(d.total_bytes / 1048576)); logln(buf);
        print_try(d.free_bytes, d.sm_count, d.cc_major, d.cc_minor);
        logln("---------------");
    }
}

// ------------------------------------------------------------------------------------------ files
const char* kGxHex = "79be667ef9dcbbac55a06295ce870b07029bfcdb2dce28d959f2815b16f81798";
std::string htcpu_name(uint64_t w, uint32_t htsz) { return std::string(kGxHex) + "_" + std::to_string(w) + "_" + std::to_string(1ull << htsz) + "_htCPUv0.BIN"; }
std::string htgpu_name(uint64_t w, uint32_t htsz) { return std::string(kGxHex) + "_" + std::to_string(w) + "_" + std::to_string(1ull << htsz) + "_htGPUv0.BIN"; }
std::string g2_name(const Options& o) { return std::to_string(o.t) + "_" + std::to_string(o.b) + "_" + std::to_string(o.p) + "_" + std::to_string(o.w) + "_g2.BIN"; }
long long file_size(const std::string& f) { FILE* fp = fopen(f.c_str(), "rb"); if (!fp) return -1; fseek(fp, 0, SEEK_END); long long n = ftell(fp); fclose(fp); return n; }
bool read_file(const std::string& f, std::vector<uint8_t>& buf) {
    long long n = file_size(f); if (n < 0) return false;
    buf.resize((size_t)n); FILE* fp = fopen(f.c_str(), "rb"); if (!fp) return false;
    size_t got = fread(buf.data(), 1, (size_t)n, fp); fclose(fp); return got == (size_t)n;
}
bool write_file(const std::string& f, const void* p, size_t n) {
    std::string tmp = f + ".tmp"; FILE* fp = fopen(tmp.c_str(), "wb"); if (!fp) return false;
    size_t put = fwrite(p, 1, n, fp); fclose(fp);
    return put == n && rename(tmp.c_str(), f.c_str()) == 0;
}
void ck(int rc, const char* what) { if (rc != BSGS_OK) { fprintf(stderr, "%s: %s (%d)\n", what, bsgs_strerror(rc), rc); exit(1); } }

void load_or_build_tables(const Options& o, std::vector<bsgs_ctx*>& ctx) {
    auto t0 = std::chrono::steady_clock::now();
    const size_t H = (size_t)1 << o.htsz, cpu_bytes = H * 4 + 4 + (size_t)o.w * 8, gpu_bytes = H * 4 + 4 + (size_t)o.w * 4;
    const size_t g2_bytes = (size_t)o.t * o.b * o.p * 64;
    const std::string f_cpu = htcpu_name(o.w, o.htsz), f_gpu = htgpu_name(o.w, o.htsz), f_g2 = g2_name(o);
    if (file_size(f_cpu) == (long long)cpu_bytes && file_size(f_gpu) == (long long)gpu_bytes) {
        logln("Load hashtable from " + f_gpu);
        std::vector<uint8_t> img;
        if (!read_file(f_cpu, img)) { fprintf(stderr, "cannot read %s\n", f_cpu.c_str()); exit(1); }
        ck(bsgs_upload_htcpu(ctx[0], img.data(), img.size()), "bsgs_upload_htcpu");      // tags + payloads: device-side verification
    } else {
        logln("Generate " + std::to_string(o.w) + " baby points on GPU#" + std::to_string(o.devices[0]));
        ck(bsgs_ht_generate(ctx[0]), "bsgs_ht_generate");
        if (!o.nosave) {
            std::vector<uint8_t> img(cpu_bytes);
            ck(bsgs_ht_export(ctx[0], 1, img.data(), img.size()), "bsgs_ht_export(cpu)"); write_file(f_cpu, img.data(), img.size());
            img.resize(gpu_bytes);
            ck(bsgs_ht_export(ctx[0], 0, img.data(), img.size()), "bsgs_ht_export(gpu)"); write_file(f_gpu, img.data(), img.size());
        }
    }
    if (ctx.size() > 1) ck(bsgs_bcast_ht(ctx.data(), (int)ctx.size()), "bsgs_bcast_ht");
    for (auto c : ctx) ck(bsgs_set_option(c, "verify", 1), "verify");
    if (file_size(f_g2) == (long long)g2_bytes) {
        logln("Load giants from " + f_g2);
        std::vector<uint8_t> g2;
        if (!read_file(f_g2, g2)) { fprintf(stderr, "cannot read %s\n", f_g2.c_str()); exit(1); }
        for (auto c : ctx) ck(bsgs_upload_giants(c, g2.data(), g2.size()), "bsgs_upload_giants");
    } else {
        logln("Generate " + std::to_string((unsigned long long)o.t * o.b * o.p) + " giant points on GPU");
        for (auto c : ctx) ck(bsgs_giants_generate(c), "bsgs_giants_generate");
        if (!o.nosave) { std::vector<uint8_t> g2(g2_bytes); ck(bsgs_giants_export(ctx[0], g2.data(), g2.size()), "bsgs_giants_export"); write_file(f_g2, g2.data(), g2.size()); }
    }
    char buf[64]; snprintf(buf, sizeof buf, "Tables ready in %.1fs", std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count()); logln(buf);
}

// ------------------------------------------------------------------------------------------ the search (main set-up 4689-4765, 5042-5064; checkerThread 3933-4296)
struct Search {
    uint64_t N, w; uint32_t p; int rank, world;
    U256 pk, stride, rank_stride, key0, center; Pt target, findpub, P0, D;
    Search(const Options& o, const Pt& tgt, const U256& start_key, int rank_, int world_) : N((uint64_t)o.t * o.b * o.p), w(o.w), p(o.p), rank(rank_), world(world_), pk(o.pk), target(tgt) {
        u128 s = (u128)4 * N * w;                                          // PRKADDBIG = 4*N*w
        stride.v[0] = (uint64_t)s; stride.v[1] = (uint64_t)(s >> 64);
        center = from_u64((uint64_t)p * w);                                // CenterBig = p*w  (p < 2^31, w < 2^32)
        findpub = pt_sub(tgt, pt_mul(hec::kG, pk));                        // FINDPUBG = target - pk*G
        key0 = addmod(mod_n(start_key), mod_n(mul_small(stride, (uint64_t)rank)));
        P0 = pt_sub(pt_sub(findpub, pt_mul(hec::kG, key0)), pt_mul(hec::kG, center));      // GlobPub, 5054-5063
        rank_stride = mul_small(stride, (uint64_t)world);
        Pt sg = pt_mul(hec::kG, rank_stride); hec::pt_neg(D, sg);          // PUBADDBIG (times the number of ranks)
    }
    U256 window_key(uint64_t j) const { return addmod(key0, mod_n(mul_small(rank_stride, j))); }
    Pt window_point(uint64_t j) const { if (!j) return P0; Pt r, s = pt_mul(D, from_u64(j)); hec::pt_add(r, P0, s); return r; }
    // The reference tests key_j > pke - pk AFTER the launch of window j (1_9_7File.pb:2512-2518): every window
    // with key_j <= width is launched and then one more, which holds the tail of the range (window j covers
    // key_j + p*w +- (2Nw + w) only).
    uint64_t windows_needed(const U256& pke, const U256& start_key) const {
        U256 width; sub_raw(width, pke, pk);
        uint64_t total = 1;
        if (cmp(width, start_key) >= 0) {
            U256 d; sub_raw(d, width, start_key);
            // floor(d / stride): stride < 2^66 and d / stride < 2^64 for every range this program accepts in practice
            long double q = 0; for (int i = 3; i >= 0; i--) q = q * 18446744073709551616.0L + (long double)d.v[i];
            long double sd = (long double)stride.v[1] * 18446744073709551616.0L + (long double)stride.v[0];
            if (q / sd >= 9.0e18L) throw UsageError{"Range is too wide for these -t/-b/-p/-w: more than 2^63 windows"};
            uint64_t est = (uint64_t)(q / sd);
            // exact correction around the floating-point estimate
            while (est > 0 && cmp(mul_small(stride, est), d) > 0) est--;
            while (cmp(mul_small(stride, est + 1), d) <= 0) est++;
            total = est + 2;
        }
        return total > (uint64_t)rank ? (total - rank + world - 1) / world : 0;
    }
    bool try_key(U256 K, U256& fin) const {                                // 1_9_7File.pb:4135-4144 and the final check against the real pubkey
        K = mod_n(K);
        if (is_zero(K)) return false;
        Pt t = pt_mul(hec::kG, K);
        if (!hec::eq(t.x, findpub.x)) return false;
        if (!hec::eq(t.y, findpub.y)) K = negmod(K);
        fin = addmod(mod_n(pk), K);
        t = pt_mul(hec::kG, fin);
        if (!hec::eq(t.x, target.x)) return false;
        if (!hec::eq(t.y, target.y)) fin = negmod(fin);
        return true;
    }
    // a record that survived device-side verification: the baby-step index b = payload is known
    bool check_record(uint32_t code, uint32_t idx, uint64_t j, uint32_t payload, U256& fin) const {
        U256 base = addmod(window_key(j), center);
        U256 giant = mod_n(mul_small(from_u64((uint64_t)(idx) + 1), 2 * w));
        if (code == 4 || payload == 0xFFFFFFFEu) return try_key(addmod(base, giant), fin) || try_key(submod(base, giant), fin);
        U256 b = from_u64((uint64_t)payload + 1);
        if (code == 5) return try_key(addmod(base, b), fin) || try_key(submod(base, b), fin);
        if (code == 1) return try_key(addmod(base, addmod(giant, b)), fin) || try_key(addmod(base, submod(giant, b)), fin);
        return try_key(submod(base, addmod(giant, b)), fin) || try_key(submod(base, submod(giant, b)), fin);
    }
};

void save_checkpoint(int listpos, const std::string& pub_full_hex, const U256& cnt, const std::string& fp) {    // 1_9_7File.pb:3913-3921
    std::string body = std::to_string(listpos) + "\n" + pub_full_hex + "\n" + hex64(cnt.v) + "\n" + fp + "\n";
    FILE* f = fopen("currentwork.temp", "w"); if (!f) return;
    fputs(body.c_str(), f); fclose(f); rename("currentwork.temp", "currentwork.txt");
}
std::string fmt_hms(double s) { int t = (int)s; char b[32]; snprintf(b, sizeof b, "%02d:%02d:%02d", (t / 3600) % 24, (t / 60) % 60, t % 60); return b; }
// the status line of the main loop (1_9_7File.pb:5119-5141): counter, workers alive, giant steps/s of every
// GPU and in total in units of 2^20, the keys each step covers (2w), job time and total time
std::string format_status(const std::string& cnt_hex, int running, const std::vector<unsigned long long>& rates, uint64_t w, double job_s, double total_s) {
    unsigned long long total = 0; std::string perf;
    for (auto r : rates) { total += r; perf += std::to_string(r >> 20) + " "; }
    const double wald = log2(2.0 * (double)w), hashd = total ? log2((double)total) : 0.0;
    char buf[128]; snprintf(buf, sizeof buf, "] = %llu MKeys/s x2^%.1f=2^%.2f Jt:", total >> 20, wald, hashd + wald);
    return "Cnt:" + cnt_hex + " [" + std::to_string(running) + "][ " + perf + buf + fmt_hms(job_s) + " Tt:" + fmt_hms(total_s);
}

int run(const std::vector<std::string>& argv) {
    const auto t_program = std::chrono::steady_clock::now();             // begintime (Tt:), 1_9_7File.pb:5135
    Options o;
    std::vector<std::string> pubs;
    try {
        o = parse_args(argv);
        for (auto& wmsg : validate(o)) logln(wmsg);
        if (!o.infile.empty()) {
            std::ifstream in(o.infile);
            if (!in) throw UsageError{"File[" + o.infile + "] not found!"};
            std::string l; while (std::getline(in, l)) { l = cut_hex(l); if (!l.empty()) pubs.push_back(l); }
        } else pubs.push_back(o.pb);
    } catch (const UsageError& e) { logln(e.msg); return 2; }
    const std::string fp = fingerprint(o);
    if (o.fingerprint_only) { logln(fp); return 0; }
    list_devices();
    logln("Current config hash[" + fp + "]");
    bool have_rec = false; int rec_pos = 0; std::string rec_pub; U256 rec_cnt;
    if (!o.wl.empty()) {                                                   // -wl, 1_9_7File.pb:4640-4688
        std::ifstream in(o.wl);
        if (!in) { logln("Recovery file [" + o.wl + "] does not exist"); return 2; }
        std::string l[4]; for (auto& s : l) { if (!std::getline(in, s) || cut_hex(s).empty()) { logln("Empty line in recovery file"); return 2; } s = cut_hex(s); }
        rec_pos = atoi(l[0].c_str()); rec_pub = l[1]; parse_hex(l[2], rec_cnt);
        if (l[3] != fp) { logln("Current and recovery configuration are different"); return 2; }
        have_rec = true;
    }
    if (o.devices.empty()) o.devices.push_back(0);
    std::vector<bsgs_ctx*> ctx(o.devices.size(), nullptr);
    for (size_t i = 0; i < ctx.size(); i++) ck(bsgs_create(&ctx[i], o.devices[i], o.t, o.b, o.p, o.w, o.htsz), "bsgs_create");
    const int R = (int)ctx.size();
    if (o.onlygen) o.nosave = false;
    load_or_build_tables(o, ctx);
    if (o.onlygen) {                                                       // onlygen1_9_6File.pb:3182-3205
        for (auto c : ctx) bsgs_destroy(c);
        logln("finished ok");
        return 0;
    }
    int found_total = 0, rc = 0;
    for (size_t li = 0; li < pubs.size() && rc == 0; li++) {
        const int listpos = (int)li + 1;
        Pt target;
        if (!parse_pubkey(pubs[li], target)) { logln("unsupported public key format"); rc = 2; break; }
        const std::string full = hex64(target.x.v) + hex64(target.y.v);
        if (have_rec && listpos != rec_pos) continue;                       // 1_9_7File.pb:5020-5024
        if (have_rec && rec_pub != full && rec_pub != pubs[li]) { logln("Find position but the keys are different"); rc = 2; break; }
        const U256 start_key = have_rec ? rec_cnt : from_u64(1);
        have_rec = false;
        logln(""); logln("Findpubkey  : " + compress(target));
        std::vector<Search> searches;
        for (int r = 0; r < R; r++) searches.emplace_back(o, target, start_key, r, R);
        logln("Searchpubkey: " + compress(searches[0].findpub));
        std::mutex mu; std::atomic<bool> stop{false}; bool have_key = false; U256 key;
        std::vector<uint64_t> done(R, 0); std::vector<unsigned long long> steps(R, 0);
        std::vector<uint64_t> nwins(R, 0);
        try { for (int r = 0; r < R; r++) nwins[r] = searches[r].windows_needed(o.pke, start_key); }
        catch (const UsageError& e) { logln(e.msg); rc = 2; break; }
        const auto t_begin = std::chrono::steady_clock::now();
        auto worker = [&](int r) {
            const Search& s = searches[r];
            const uint64_t nwin = nwins[r];
            std::vector<bsgs_hit> hits(4096);
            uint64_t d0 = 0;
            while (d0 < nwin && !stop) {
                const uint32_t chunk = (uint32_t)std::min<uint64_t>(nwin - d0, 128);      // 8 launches of 16 windows, two in flight inside bsgs_run
                Pt P = s.window_point(d0);
                uint8_t p0[64], dd[64]; memcpy(p0, P.x.v, 32); memcpy(p0 + 32, P.y.v, 32); memcpy(dd, s.D.x.v, 32); memcpy(dd + 32, s.D.y.v, 32);
                uint32_t nh = 0; uint64_t wd = 0;
                int e = bsgs_run(ctx[r], p0, dd, chunk, std::min<uint32_t>(chunk, 16), 1, hits.data(), (uint32_t)hits.size(), &nh, &wd);
                if (e != BSGS_OK && e != BSGS_E_OVERFLOW) { fprintf(stderr, "bsgs_run: %s\n", bsgs_strerror(e)); stop = true; return; }
                if (e == BSGS_E_OVERFLOW && hits.size() < (1u << 22)) { hits.resize(hits.size() * 4); continue; }      // rerun this chunk with a bigger record buffer
                nh = std::min<uint32_t>(nh, (uint32_t)hits.size());
                for (uint32_t i = 0; i < nh; i++) {
                    U256 fin;
                    if (hits[i].reserved != 0xFFFFFFFFu && s.check_record(hits[i].code, hits[i].idx, d0 + hits[i].window, hits[i].reserved, fin)) {
                        std::lock_guard<std::mutex> g(mu); key = fin; have_key = true; stop = true;
                    }
                }
                d0 += wd;
                std::lock_guard<std::mutex> g(mu); done[r] = d0; steps[r] += 2ull * s.N * wd;
            }
        };
        std::vector<std::thread> th;
        for (int r = 0; r < R; r++) th.emplace_back(worker, r);
        std::atomic<int> alive{R};
        // status line + checkpoint from the main thread (1_9_7File.pb:5119-5141, 3897-3931)
        std::thread joiner([&] { for (auto& t : th) { t.join(); alive--; } });
        auto last_log = std::chrono::steady_clock::now(), last_save = last_log; std::vector<unsigned long long> last_steps(R, 0);
        while (alive > 0) {
            std::this_thread::sleep_for(std::chrono::milliseconds(2));
            auto now = std::chrono::steady_clock::now();
            auto min_cnt = [&] { std::lock_guard<std::mutex> g(mu); U256 m = searches[0].window_key(done[0]); for (int r = 1; r < R; r++) { U256 k = searches[r].window_key(done[r]); if (cmp(k, m) < 0) m = k; } return m; };
            if (std::chrono::duration<double>(now - last_log).count() > 2) {
                std::vector<unsigned long long> st; { std::lock_guard<std::mutex> g(mu); st = steps; }
                const double dt = std::chrono::duration<double>(now - last_log).count();
                std::vector<unsigned long long> rates(R);
                for (int r = 0; r < R; r++) rates[r] = (unsigned long long)((double)(st[r] - last_steps[r]) / dt);
                last_steps = st; last_log = now;
                logln(format_status(hex_short(min_cnt()), (int)alive, rates, o.w, std::chrono::duration<double>(now - t_begin).count(),
                                    std::chrono::duration<double>(now - t_program).count()));
            }
            if (std::chrono::duration<double>(now - last_save).count() >= o.wt) { save_checkpoint(listpos, full, min_cnt(), fp); last_save = now; }
        }
        joiner.join();
        if (have_key) {
            const std::string tag = "KEY[" + std::to_string(listpos) + "]: ";
            std::string l1 = tag + "0x" + hex64(key.v), l2 = std::string(tag.size() - 5, ' ') + "Pub: " + compress(target);
            logln(l1); logln(l2);
            FILE* f = fopen("win.txt", "a"); if (f) { fprintf(f, "%s\n%s\n", l1.c_str(), l2.c_str()); fclose(f); }     // #WINFILE, 1_9_7File.pb:19, 5151-5157
            logln("Job time " + fmt_hms(std::chrono::duration<double>(std::chrono::steady_clock::now() - t_begin).count()) + "s");
            found_total++;
        } else logln("Reached end of space");
    }
    if (rc == 0) logln("finished, found " + std::to_string(found_total) + " of " + std::to_string(pubs.size()));
    for (auto c : ctx) bsgs_destroy(c);
    return rc;
}

}  // namespace

int main(int argc, char** argv) {
    std::vector<std::string> args(argv + 1, argv + argc);
    return run(args);
}
