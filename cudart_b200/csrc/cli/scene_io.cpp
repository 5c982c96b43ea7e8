// cudart_b200/csrc/cli/scene_io.cpp - see scene_io.h.  No CUDA here; BVH and light list come through the C ABI.
#include "scene_io.h"

#include <zlib.h>

#include <algorithm>
#include <array>
#include <atomic>
#include <cerrno>
#include <cfloat>
#include <charconv>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <memory>
#include <sstream>
#include <thread>

#include "../../../include/crt_abi.h"

namespace crtcli {
namespace {

[[noreturn]] void die(const std::string& what) { throw std::runtime_error(what); }

// ---- line helpers with the reference parser's token rules (obj_load.h:330-402) -------------------
std::string first_token(const std::string& in) {
    const size_t a = in.find_first_not_of(" \t");
    if (in.empty() || a == std::string::npos) return "";
    const size_t b = in.find_first_of(" \t", a);
    return b == std::string::npos ? in.substr(a) : in.substr(a, b - a);
}
std::string tail_of(const std::string& in) {
    const size_t a = in.find_first_not_of(" \t");
    const size_t sp = in.find_first_of(" \t", a);
    const size_t b = in.find_first_not_of(" \t", sp);
    const size_t e = in.find_last_not_of(" \t");
    if (b == std::string::npos) return "";
    return e == std::string::npos ? in.substr(b) : in.substr(b, e - b + 1);
}
// Splits on a one-character separator; an empty field between two separators is kept (that is how "1//3" yields
// an empty texture index), a separator right after a non-empty field is consumed.
std::vector<std::string> split_on(const std::string& in, char sep) {
    std::vector<std::string> out;
    std::string cur;
    for (size_t i = 0; i < in.size(); i++) {
        if (in[i] == sep) {
            if (!cur.empty()) { out.push_back(cur); cur.clear(); }
            else out.push_back("");
        } else if (i + 1 >= in.size()) {
            cur += in[i];
            out.push_back(cur);
            break;
        } else {
            cur += in[i];
        }
    }
    return out;
}
template <class T>
const T& element_at(const std::vector<T>& v, const std::string& index) {      // 1-based, negative = from the end
    int idx = std::stoi(index);
    idx = idx < 0 ? (int)v.size() + idx : idx - 1;
    if (idx < 0 || idx >= (int)v.size()) die("OBJ index out of range: " + index);
    return v[(size_t)idx];
}

bool file_exists(const std::string& p) { return std::ifstream(p.c_str()).good(); }

// ---- a small JSON reader (the camera file is one flat object of numbers, strings and number arrays) -----
struct Json {
    enum Kind { Null, Num, Str, Arr, Obj, Bool } kind = Null;
    double num = 0;
    std::string str;
    std::vector<Json> arr;
    std::vector<std::pair<std::string, Json>> obj;
    const Json& at(const std::string& k) const {
        for (auto& kv : obj) if (kv.first == k) return kv.second;
        die("camera JSON: missing key \"" + k + "\"");
    }
};
struct JsonParser {
    const std::string& s;
    size_t i = 0;
    void ws() { while (i < s.size() && std::isspace((unsigned char)s[i])) i++; }
    Json value() {
        ws();
        if (i >= s.size()) die("camera JSON: unexpected end");
        Json j;
        const char c = s[i];
        if (c == '{') {
            j.kind = Json::Obj; i++; ws();
            if (s[i] == '}') { i++; return j; }
            for (;;) {
                ws();
                Json k = value();
                if (k.kind != Json::Str) die("camera JSON: object key must be a string");
                ws();
                if (s[i++] != ':') die("camera JSON: ':' expected");
                j.obj.emplace_back(k.str, value());
                ws();
                if (s[i] == ',') { i++; continue; }
                if (s[i] == '}') { i++; return j; }
                die("camera JSON: ',' or '}' expected");
            }
        }
        if (c == '[') {
            j.kind = Json::Arr; i++; ws();
            if (s[i] == ']') { i++; return j; }
            for (;;) {
                j.arr.push_back(value());
                ws();
                if (s[i] == ',') { i++; continue; }
                if (s[i] == ']') { i++; return j; }
                die("camera JSON: ',' or ']' expected");
            }
        }
        if (c == '"') {
            j.kind = Json::Str; i++;
            while (i < s.size() && s[i] != '"') { if (s[i] == '\\' && i + 1 < s.size()) i++; j.str += s[i++]; }
            i++;
            return j;
        }
        if (!s.compare(i, 4, "true")) { j.kind = Json::Bool; j.num = 1; i += 4; return j; }
        if (!s.compare(i, 5, "false")) { j.kind = Json::Bool; i += 5; return j; }
        if (!s.compare(i, 4, "null")) { i += 4; return j; }
        char* end = nullptr;
        j.kind = Json::Num;
        j.num = std::strtod(s.c_str() + i, &end);           // picojson parses numbers with strtod as well
        if (end == s.c_str() + i) die("camera JSON: bad value");
        i = (size_t)(end - s.c_str());
        return j;
    }
};

float len3(const crt::F3& v) { return sqrtf(v.x * v.x + v.y * v.y + v.z * v.z); }
crt::F3 norm3(const crt::F3& v) { const float n = len3(v); return {v.x / n, v.y / n, v.z / n}; }
crt::F3 cross3(const crt::F3& a, const crt::F3& b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
crt::F3 vec_from(const Json& a) {
    if (a.kind != Json::Arr || a.arr.size() < 3) die("camera JSON: 3-vector expected");
    return {(float)a.arr[0].num, (float)a.arr[1].num, (float)a.arr[2].num};
}

}  // namespace

// ---- .env (env_loader.cpp:14-37) ------------------------------------------------------------------------
EnvFile::EnvFile(const std::string& path) {
    std::ifstream f(path);
    if (!f.is_open()) die("Error opening " + path);
    const std::string keys[3] = {"MESHES_PATH", "CAMERA_PATH", "TEXTURES_PATH"};
    std::string* slots[3] = {&meshes, &camera, &textures};
    std::string line;
    while (std::getline(f, line)) {
        if (!line.empty() && (line[0] == '#' || line[0] == '\n')) continue;
        for (int k = 0; k < 3; k++)
            if (line.find(keys[k]) != std::string::npos) {      // the value is whatever follows KEY plus one character
                *slots[k] = line.size() > keys[k].size() + 1 ? line.substr(keys[k].size() + 1) : "";
                break;
            }
    }
}

// ---- camera (json_loader.cpp:25-47) ---------------------------------------------------------------------
crt::RefCamera load_camera_json(const std::string& path, unsigned width, unsigned height) {
    if (!file_exists(path)) die("File " + path + " not found.");
    std::ifstream f(path);
    std::stringstream ss;
    ss << f.rdbuf();
    const std::string text = ss.str();
    JsonParser p{text};
    const Json root = p.value();
    if (root.kind != Json::Obj) die("camera JSON: object expected loading camera!");
    const float fov_deg = (float)root.at("fieldOfView").num;
    const float focal = (float)root.at("focalLength").num;
    const float fstop = (float)root.at("fStop").num;
    crt::RefCamera c{};
    c.xpixels = (int32_t)width;
    c.ypixels = (int32_t)height;
    c.fov = tanf(fov_deg * 0.5f * M_PI / 180.0f);
    c.eye = vec_from(root.at("eye"));
    const crt::F3 dir = vec_from(root.at("viewDirection"));
    c.focusDistance = len3(dir);
    c.dir = norm3(dir);
    c.up = norm3(vec_from(root.at("upDirection")));
    c.right = norm3(cross3(c.dir, c.up));
    c.apertureWidth = focal / fstop;
    c.aspect = (float)width / (float)height;
    return c;
}

// ---- OBJ / MTL (obj_load.h:441-723, 1016-1186) ------------------------------------------------------------
namespace {
void load_mtl(const std::string& path, std::vector<ObjMaterial>& out) {
    if (path.size() < 4 || path.substr(path.size() - 4) != ".mtl") return;
    std::ifstream f(path);
    if (!f.is_open()) return;
    ObjMaterial cur;
    bool open = false;
    std::string line;
    auto triple = [](const std::string& l, float* dst) {
        const std::vector<std::string> t = split_on(tail_of(l), ' ');
        if (t.size() != 3) return;
        for (int k = 0; k < 3; k++) dst[k] = std::stof(t[k]);
    };
    while (std::getline(f, line)) {
        const std::string key = first_token(line);
        if (key == "newmtl") {
            if (open) { out.push_back(cur); cur = ObjMaterial(); }
            open = true;
            cur.name = line.size() > 7 ? tail_of(line) : "none";
        }
        if (key == "Ka") triple(line, cur.Ka);
        if (key == "Kd") triple(line, cur.Kd);
        if (key == "Ks") triple(line, cur.Ks);
        if (key == "Ns") cur.Ns = std::stof(tail_of(line));
        if (key == "Ni") cur.Ni = std::stof(tail_of(line));
        if (key == "map_Kd") cur.map_Kd = tail_of(line);
        if (key == "type") cur.type = tail_of(line);                  // CudaRT-specific keys (obj_load.h:1161-1170)
        if (key == "diffuse") cur.diffuse = std::stof(tail_of(line));
    }
    out.push_back(cur);
}
}  // namespace

namespace {
// ---- fast path of load_obj ----------------------------------------------------------------------------------------
// A 1M-triangle OBJ has 1.4 M lines; going through std::getline + token vectors + std::stof / std::stoi for each of them took
// 1.7 s of the command line's 2.5 s before the first sample.  The lines that make up such a file have one strict shape -
// `v x y z`, `vn x y z`, `vt u v`, `f a b c` with corners a, a/b, a//c or a/b/c, single blanks, nothing else - and are parsed
// here straight from the file buffer.  Anything that deviates in any way (leading or doubled blanks, tabs, a trailing blank
// or '\r', other keys, numbers that are not plain decimals, indices of more than 9 digits, out-of-range values) is NOT
// handled: the line goes through line_slow(), the reference parser's own token rules, so quirks and error behaviour stay
// what they were.

// [sign] digits [. digits] [e [sign] digits] with at least one digit in the mantissa; returns the end or nullptr
const char* plain_decimal_end(const char* p) {
    if (*p == '-' || *p == '+') p++;
    const char* d = p;
    while (*p >= '0' && *p <= '9') p++;
    size_t digits = (size_t)(p - d);
    if (*p == '.') {
        p++;
        d = p;
        while (*p >= '0' && *p <= '9') p++;
        digits += (size_t)(p - d);
    }
    if (digits == 0) return nullptr;
    if (*p == 'e' || *p == 'E') {
        const char* q = p + 1;
        if (*q == '-' || *q == '+') q++;
        if (!(*q >= '0' && *q <= '9')) return nullptr;
        while (*q >= '0' && *q <= '9') q++;
        p = q;
    }
    return p;
}
// n plain decimals separated by single blanks, ending exactly at `end`; the values std::stof would return (stof is strtof
// plus error checks: out-of-range results make the caller fall back to it)
bool parse_floats(const char* p, const char* end, float* dst, int n) {
    for (int k = 0; k < n; k++) {
        const char* e = plain_decimal_end(p);
        if (!e || e > end) return false;
        if (k + 1 < n ? (*e != ' ') : (e != end)) return false;
        // std::from_chars and strtof are both correctly rounded, so they agree wherever the result is an ordinary number
        // (checked on 2e7 random decimals); zeros, subnormals and overflows take strtof, whose ERANGE is stof's exception
        float v = 0.f;
        const std::from_chars_result r = std::from_chars(p, e, v);
        if (r.ec != std::errc() || r.ptr != e || !(std::fabs(v) >= FLT_MIN) || !std::isfinite(v)) {
            errno = 0;
            char* stop = nullptr;
            v = std::strtof(p, &stop);
            if (stop != e || errno == ERANGE) return false;
        }
        dst[k] = v;
        p = e + 1;
    }
    return true;
}
// [-] 1..9 digits: what std::stoi returns for such a token
const char* parse_index(const char* p, int& out) {
    const bool neg = *p == '-';
    if (neg) p++;
    const char* d = p;
    int v = 0;
    while (*p >= '0' && *p <= '9' && p - d < 10) v = v * 10 + (*p++ - '0');
    if (p == d || p - d > 9) return nullptr;
    out = neg ? -v : v;
    return p;
}
template <class T>
const T* element_or_null(const std::vector<T>& v, int idx) {          // element_at without the throw
    idx = idx < 0 ? (int)v.size() + idx : idx - 1;
    return (idx < 0 || idx >= (int)v.size()) ? nullptr : &v[(size_t)idx];
}
}  // namespace

ObjFile load_obj(const std::string& path) {
    ObjFile out;
    if (path.size() < 4 || path.substr(path.size() - 4) != ".obj") die("Failed to load mesh for " + path);
    std::string text;
    {
        std::ifstream f(path, std::ios::binary);
        if (!f.is_open()) die("Failed to load mesh for " + path);
        f.seekg(0, std::ios::end);
        const std::streamoff size = f.tellg();
        f.seekg(0, std::ios::beg);
        text.resize(size > 0 ? (size_t)size : 0);
        if (size > 0) f.read(&text[0], size);
        text.resize((size_t)std::max<std::streamsize>(f.gcount(), 0));
    }
    std::vector<std::array<float, 3>> pos, nor;
    std::vector<std::array<float, 2>> tex;
    std::vector<std::string> mesh_mat_names;
    std::vector<ObjVertex> cur;
    std::string meshname, line;
    bool listening = false;
    {   // three vertices per `f` line: size the vertex soup once instead of growing 96 MB arrays by doubling
        size_t faces = 0;
        for (size_t i = 0; i + 1 < text.size(); i++)
            if (text[i] == 'f' && text[i + 1] == ' ' && (i == 0 || text[i - 1] == '\n')) faces++;
        out.all_vertices.reserve(3 * faces);
        cur.reserve(3 * faces);
    }
    auto close_mesh = [&](const std::string& name) {
        ObjMesh m;
        m.name = name;
        const size_t left = cur.capacity() - cur.size();
        m.vertices.swap(cur);
        cur.reserve(left);
        out.meshes.push_back(std::move(m));
    };
    // the reference parser's rules, line by line (obj_load.h:441-723)
    auto line_slow = [&](const std::string& line) {
        const std::string key = first_token(line);
        if (key == "o" || key == "g" || (!line.empty() && line[0] == 'g')) {
            const std::string name = (key == "o" || key == "g") ? tail_of(line) : "unnamed";
            if (!listening) listening = true;
            else if (!cur.empty()) close_mesh(meshname);
            meshname = name;
        }
        if (key == "v" || key == "vn") {
            const std::vector<std::string> t = split_on(tail_of(line), ' ');
            if (t.size() < 3) die("OBJ: malformed line: " + line);
            (key == "v" ? pos : nor).push_back({std::stof(t[0]), std::stof(t[1]), std::stof(t[2])});
        }
        if (key == "vt") {
            const std::vector<std::string> t = split_on(tail_of(line), ' ');
            if (t.size() < 2) die("OBJ: malformed line: " + line);
            tex.push_back({std::stof(t[0]), std::stof(t[1])});
        }
        if (key == "f") {
            const std::vector<std::string> corners = split_on(tail_of(line), ' ');
            if (corners.size() != 3)
                die("OBJ: only triangulated faces are supported (the reference counts three vertices per face, scene.h:27-29): " + line);
            ObjVertex v[3];
            bool no_normal = false;
            for (int k = 0; k < 3; k++) {
                const std::vector<std::string> idx = split_on(corners[k], '/');
                const auto& p = element_at(pos, idx.at(0));
                std::memcpy(v[k].p, p.data(), 12);
                if (idx.size() >= 2 && !idx[1].empty()) { const auto& t = element_at(tex, idx[1]); v[k].t[0] = t[0]; v[k].t[1] = t[1]; }
                if (idx.size() == 3) { const auto& n = element_at(nor, idx[2]); std::memcpy(v[k].n, n.data(), 12); }
                else no_normal = true;
            }
            if (no_normal) {          // face normal from the corner positions (obj_load.h:836-847)
                const float a[3] = {v[0].p[0] - v[1].p[0], v[0].p[1] - v[1].p[1], v[0].p[2] - v[1].p[2]};
                const float b[3] = {v[2].p[0] - v[1].p[0], v[2].p[1] - v[1].p[1], v[2].p[2] - v[1].p[2]};
                const float n[3] = {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
                for (auto& vv : v) std::memcpy(vv.n, n, 12);
            }
            for (auto& vv : v) { cur.push_back(vv); out.all_vertices.push_back(vv); }
        }
        if (key == "usemtl") {
            mesh_mat_names.push_back(tail_of(line));
            if (!cur.empty()) close_mesh(meshname + "_2");          // material change inside a group starts a new mesh
        }
        if (key == "mtllib") {
            const size_t slash = path.find_last_of('/');
            load_mtl((slash == std::string::npos ? std::string() : path.substr(0, slash + 1)) + tail_of(line), out.materials);
        }
    };
    // the same four line kinds in their strict form, straight from the buffer; false = not handled, nothing changed
    auto line_fast = [&](const char* b, const char* e) -> bool {
        if (e - b < 3) return false;
        float x[3];
        if (b[0] == 'v' && b[1] == ' ') {
            if (!parse_floats(b + 2, e, x, 3)) return false;
            pos.push_back({x[0], x[1], x[2]});
            return true;
        }
        if (b[0] == 'v' && b[1] == 'n' && b[2] == ' ') {
            if (!parse_floats(b + 3, e, x, 3)) return false;
            nor.push_back({x[0], x[1], x[2]});
            return true;
        }
        if (b[0] == 'v' && b[1] == 't' && b[2] == ' ') {
            if (!parse_floats(b + 3, e, x, 2)) return false;
            tex.push_back({x[0], x[1]});
            return true;
        }
        if (b[0] == 'f' && b[1] == ' ') {
            ObjVertex v[3];
            bool no_normal = false;
            const char* p = b + 2;
            for (int k = 0; k < 3; k++) {
                int ip = 0, it = 0, in = 0;
                bool has_t = false, has_n = false;
                if (!(p = parse_index(p, ip))) return false;
                if (*p == '/') {
                    p++;
                    if (*p == '/') {                                  // a//c
                        p++;
                        if (!(p = parse_index(p, in))) return false;
                        has_n = true;
                    } else {                                          // a/b or a/b/c
                        if (!(p = parse_index(p, it))) return false;
                        has_t = true;
                        if (*p == '/') {
                            p++;
                            if (!(p = parse_index(p, in))) return false;
                            has_n = true;
                        }
                    }
                }
                if (k < 2 ? (*p != ' ') : (p != e)) return false;
                p++;
                const auto* pp = element_or_null(pos, ip);
                const auto* tp = has_t ? element_or_null(tex, it) : nullptr;
                const auto* np = has_n ? element_or_null(nor, in) : nullptr;
                if (!pp || (has_t && !tp) || (has_n && !np)) return false;      // the slow path reports the bad index
                std::memcpy(v[k].p, pp->data(), 12);
                if (tp) { v[k].t[0] = (*tp)[0]; v[k].t[1] = (*tp)[1]; }
                if (np) std::memcpy(v[k].n, np->data(), 12);
                else no_normal = true;
            }
            if (no_normal) {          // face normal from the corner positions (obj_load.h:836-847)
                const float a[3] = {v[0].p[0] - v[1].p[0], v[0].p[1] - v[1].p[1], v[0].p[2] - v[1].p[2]};
                const float c[3] = {v[2].p[0] - v[1].p[0], v[2].p[1] - v[1].p[1], v[2].p[2] - v[1].p[2]};
                const float n[3] = {a[1] * c[2] - a[2] * c[1], a[2] * c[0] - a[0] * c[2], a[0] * c[1] - a[1] * c[0]};
                for (auto& vv : v) std::memcpy(vv.n, n, 12);
            }
            for (auto& vv : v) { cur.push_back(vv); out.all_vertices.push_back(vv); }
            return true;
        }
        return false;
    };
    // std::getline's view of the file: lines end at '\n', a last line without one still counts, no empty line after a
    // final '\n'.  (The buffer is NUL-terminated, so the number parsers above stop at the latest there.)
    const bool only_slow = getenv("CRT_OBJ_SLOW") != nullptr;      // tests: every line through the reference parser's rules
    const char* const base = text.c_str();
    const char* const file_end = base + text.size();
    for (const char* b = base; b < file_end;) {
        const char* e = (const char*)std::memchr(b, '\n', (size_t)(file_end - b));
        if (!e) e = file_end;
        if (only_slow || !line_fast(b, e)) {
            line.assign(b, e);
            line_slow(line);
        }
        b = e + 1;
    }
    if (!cur.empty()) close_mesh(meshname);
    // mesh i takes the material named by the i-th `usemtl` (obj_load.h:698-713, D25)
    for (size_t i = 0; i < mesh_mat_names.size() && i < out.meshes.size(); i++)
        for (const ObjMaterial& m : out.materials)
            if (m.name == mesh_mat_names[i]) { out.meshes[i].material = m; break; }
    if (out.meshes.empty()) die("Failed to load mesh for " + path);
    return out;
}

// ---- PNG in / out ---------------------------------------------------------------------------------------------
namespace {
uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
int paeth(int a, int b, int c) {
    const int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}
}  // namespace

Texture load_png_as_float3(const std::string& path) {
    if (!file_exists(path)) die("File " + path + " not found.");
    std::ifstream f(path, std::ios::binary);
    std::vector<uint8_t> buf((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', '\r', '\n', 0x1a, '\n'};
    if (buf.size() < 33 || std::memcmp(buf.data(), sig, 8)) die(path + ": not a PNG file");
    unsigned w = 0, h = 0, depth = 0, ctype = 0, interlace = 0;
    std::vector<uint8_t> idat, plte;
    for (size_t off = 8; off + 12 <= buf.size();) {
        const uint32_t len = be32(&buf[off]);
        const std::string tag((const char*)&buf[off + 4], 4);
        const uint8_t* data = &buf[off + 8];
        if (off + 12 + len > buf.size()) die(path + ": truncated PNG");
        if (tag == "IHDR") { w = be32(data); h = be32(data + 4); depth = data[8]; ctype = data[9]; interlace = data[12]; }
        else if (tag == "PLTE") plte.assign(data, data + len);
        else if (tag == "IDAT") idat.insert(idat.end(), data, data + len);
        else if (tag == "IEND") break;
        off += 12 + len;
    }
    const int ch = ctype == 2 ? 3 : ctype == 6 ? 4 : ctype == 3 ? 1 : 0;
    if (!w || !h || depth != 8 || !ch || interlace) die(path + ": only 8-bit non-interlaced RGB / RGBA / palette PNGs are supported");
    const size_t stride = (size_t)w * ch;
    std::vector<uint8_t> raw((stride + 1) * h);
    uLongf rawlen = (uLongf)raw.size();
    if (uncompress(raw.data(), &rawlen, idat.data(), (uLong)idat.size()) != Z_OK || rawlen != raw.size()) die(path + ": bad PNG data");
    std::vector<uint8_t> px(stride * h);
    for (unsigned y = 0; y < h; y++) {
        const uint8_t ft = raw[(stride + 1) * y];
        const uint8_t* in = &raw[(stride + 1) * y + 1];
        uint8_t* row = &px[stride * y];
        const uint8_t* up = y ? &px[stride * (y - 1)] : nullptr;
        for (size_t x = 0; x < stride; x++) {
            const int a = x >= (size_t)ch ? row[x - ch] : 0, b = up ? up[x] : 0, c = (up && x >= (size_t)ch) ? up[x - ch] : 0;
            int v = in[x];
            switch (ft) {
                case 0: break;
                case 1: v += a; break;
                case 2: v += b; break;
                case 3: v += (a + b) >> 1; break;
                case 4: v += paeth(a, b, c); break;
                default: die(path + ": bad PNG filter");
            }
            row[x] = (uint8_t)v;
        }
    }
    Texture t;
    t.w = w;
    t.h = h;
    t.rgb.resize((size_t)w * h * 3);
    // stbi_loadf on an LDR file: channel -> pow(v / 255, 2.2) (stb_image.h:1671-1690, compiled as C++: the float pow)
    float lut[256];
    for (int v = 0; v < 256; v++) lut[v] = (float)(std::pow(v / 255.0f, 2.2f) * 1.0f);
    for (size_t i = 0; i < (size_t)w * h; i++)
        for (int k = 0; k < 3; k++) {
            uint8_t v;
            if (ctype == 3) { const size_t e = (size_t)px[i] * 3 + k; v = e < plte.size() ? plte[e] : 0; }
            else v = px[i * ch + k];
            t.rgb[3 * i + k] = lut[v];
        }
    return t;
}

// One zlib stream, deflated in bands of rows on all host threads (the way pigz does it): every band is a raw deflate
// stream of its own that ends on a byte boundary (Z_FULL_FLUSH; the last one with Z_FINISH), the bands are concatenated
// behind one zlib header and the Adler-32 of the whole is combined from the bands'.  The band height is fixed by the image
// width alone, so the file does not depend on the number of threads.  A 1920x1080 frame took 0.28 s on one thread - longer
// than 64 samples per pixel take to render - which is what `--progress N` pays per refresh.
void write_png_rgba8(const std::string& path, const uint8_t* rgba, unsigned w, unsigned h) {
    const size_t stride = (size_t)w * 4 + 1;                                   // filter byte (0 = none) + the row
    const unsigned band_rows = (unsigned)std::max<size_t>(1, (256u << 10) / stride);
    const unsigned n_bands = std::max(1u, (h + band_rows - 1) / band_rows);
    std::vector<std::vector<uint8_t>> band(n_bands);
    std::vector<uLong> band_adler(n_bands, 1);
    std::vector<size_t> band_raw(n_bands, 0);
    std::atomic<unsigned> next{0};
    std::atomic<bool> failed{false};
    auto work = [&] {
        std::vector<uint8_t> raw;
        for (unsigned k; (k = next.fetch_add(1)) < n_bands;) {
            const unsigned y0 = std::min(h, k * band_rows), y1 = std::min(h, y0 + band_rows);
            raw.resize(stride * (y1 - y0));
            for (unsigned y = y0; y < y1; y++) {
                raw[stride * (y - y0)] = 0;
                std::memcpy(&raw[stride * (y - y0) + 1], rgba + (size_t)w * 4 * y, (size_t)w * 4);
            }
            z_stream zs;
            std::memset(&zs, 0, sizeof zs);
            if (deflateInit2(&zs, 6, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) { failed = true; return; }
            band[k].resize(deflateBound(&zs, (uLong)raw.size()) + 16);
            zs.next_in = raw.data(); zs.avail_in = (uInt)raw.size();
            zs.next_out = band[k].data(); zs.avail_out = (uInt)band[k].size();
            const bool last = k + 1 == n_bands;
            const int rc = deflate(&zs, last ? Z_FINISH : Z_FULL_FLUSH);
            if ((last ? rc != Z_STREAM_END : rc != Z_OK) || zs.avail_in != 0) failed = true;
            band[k].resize(band[k].size() - zs.avail_out);
            deflateEnd(&zs);
            band_adler[k] = adler32(adler32(0L, Z_NULL, 0), raw.data(), (uInt)raw.size());
            band_raw[k] = raw.size();
        }
    };
    {
        const unsigned n_threads = std::min(n_bands, crt::host_threads());
        std::vector<std::unique_ptr<crt::Worker>> th;
        for (unsigned t = 1; t < n_threads; t++) th.emplace_back(new crt::Worker(work));
        work();
        for (auto& t : th) t->join();
    }
    if (failed) die("PNG compression failed");
    std::vector<uint8_t> comp = {0x78, 0x9c};                                    // zlib header: deflate, 32 KB window, default level
    uLong adler = adler32(0L, Z_NULL, 0);
    for (unsigned k = 0; k < n_bands; k++) {
        comp.insert(comp.end(), band[k].begin(), band[k].end());
        adler = adler32_combine(adler, band_adler[k], (z_off_t)band_raw[k]);
    }
    for (int sh = 24; sh >= 0; sh -= 8) comp.push_back((uint8_t)(adler >> sh));
    const uLongf clen = (uLongf)comp.size();
    FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) die("cannot write " + path);
    auto chunk = [&](const char* tag, const uint8_t* data, uint32_t len) {
        uint8_t hdr[8] = {(uint8_t)(len >> 24), (uint8_t)(len >> 16), (uint8_t)(len >> 8), (uint8_t)len,
                          (uint8_t)tag[0], (uint8_t)tag[1], (uint8_t)tag[2], (uint8_t)tag[3]};
        uLong crc = crc32(0L, hdr + 4, 4);
        if (len) crc = crc32(crc, data, len);
        const uint8_t tailb[4] = {(uint8_t)(crc >> 24), (uint8_t)(crc >> 16), (uint8_t)(crc >> 8), (uint8_t)crc};
        std::fwrite(hdr, 1, 8, f);
        if (len) std::fwrite(data, 1, len, f);
        std::fwrite(tailb, 1, 4, f);
    };
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', '\r', '\n', 0x1a, '\n'};
    std::fwrite(sig, 1, 8, f);
    const uint8_t ihdr[13] = {(uint8_t)(w >> 24), (uint8_t)(w >> 16), (uint8_t)(w >> 8), (uint8_t)w,
                              (uint8_t)(h >> 24), (uint8_t)(h >> 16), (uint8_t)(h >> 8), (uint8_t)h, 8, 6, 0, 0, 0};
    chunk("IHDR", ihdr, 13);
    chunk("IDAT", comp.data(), (uint32_t)clen);
    chunk("IEND", nullptr, 0);
    std::fclose(f);
}

// ---- scene assembly (scene.cpp:55-203) ------------------------------------------------------------------------
namespace {
enum { DIFFUSE = 0, SPECULAR, DIFFSPEC, MICROFACET, REFRACTIVE, EMISSIVE };

struct Mat {                                         // material.h:17-55
    int bsdf = DIFFUSE;
    float kd[3] = {0, 0, 0}, ka[3] = {0, 0, 0}, ks[3] = {0, 0, 0};
    float ns = 0.f, ni = 1.f, diffuse = 1.f;
    int tex = -1;
};
// The reference's float3 operator!= is true only when ALL three components differ (linalg.h:12), so its
// Material::operator== calls two materials equal unless some colour differs in every component (D10).
bool all_components_differ(const float* a, const float* b) { return a[0] != b[0] && a[1] != b[1] && a[2] != b[2]; }
bool ref_equal(const Mat& a, const Mat& b) {
    if (a.bsdf != b.bsdf) return false;
    if (all_components_differ(a.kd, b.kd) || all_components_differ(a.ka, b.ka) || all_components_differ(a.ks, b.ks)) return false;
    return !(a.ns != b.ns) && !(a.ni != b.ni) && !(a.diffuse != b.diffuse);
}
// material.h:57-79 - not a strict weak ordering (D11); the map built with it decides the material ids.
struct RefLess {
    bool operator()(const Mat& l, const Mat& r) const {
        if (l.bsdf != r.bsdf) return l.bsdf < r.bsdf;
        if (l.kd[0] < r.kd[0] || l.kd[1] < r.kd[1] || l.kd[2] < r.kd[2]) return true;
        if (l.ka[0] < r.ka[0] || l.ka[1] < r.ka[1] || l.ka[2] < r.ka[2]) return true;
        if (l.ks[0] < r.ks[0] || l.ks[1] < r.ks[1] || l.ks[2] < r.ks[2]) return true;
        if (l.ns < r.ns) return true;
        if (l.ni < r.ni) return true;
        return false;
    }
};
Mat from_mtl(const ObjMaterial& m) {                 // scene.cpp:175-190
    static const std::map<std::string, int> kinds = {{"LAMBERT", DIFFUSE}, {"SPECULAR", SPECULAR}, {"DIFFSPEC", DIFFSPEC},
                                                     {"MICROFACET", MICROFACET}, {"REFRACTIVE", REFRACTIVE}, {"EMISSIVE", EMISSIVE}};
    Mat o;
    for (int k = 0; k < 3; k++) { o.ka[k] = m.Ka[k] * 3.0f; o.kd[k] = m.Kd[k]; o.ks[k] = m.Ks[k]; }   // LIGHTS_GAIN 3
    o.ns = m.Ns; o.ni = m.Ni; o.diffuse = m.diffuse;
    const auto it = kinds.find(m.type);
    o.bsdf = it == kinds.end() ? DIFFUSE : it->second;
    return o;
}
}  // namespace

SceneArrays assemble_scene(const ObjFile& obj, const crt::RefCamera& cam, const std::string& textures_dir) {
    SceneArrays s;
    s.camera = cam;
    // textures: every distinct map_Kd of the MTL file, in file order (scene.cpp:55-71, texture_store.cpp:37-71)
    std::vector<std::string> tex_files;
    for (const ObjMaterial& m : obj.materials)
        if (!m.map_Kd.empty() && std::find(tex_files.begin(), tex_files.end(), m.map_Kd) == tex_files.end()) tex_files.push_back(m.map_Kd);
    s.n_tex = (uint32_t)tex_files.size();
    s.tex_offsets.push_back(0);
    uint32_t total = 0, last_w = 0;
    for (const std::string& file : tex_files) {
        const Texture t = load_png_as_float3(textures_dir + file);
        std::cout << "Loading " << textures_dir + file << ": not hdr." << std::endl;
        s.tex_dims.push_back(t.w);
        s.tex_dims.push_back(t.h);
        s.tex_data.insert(s.tex_data.end(), t.rgb.begin(), t.rgb.end());
        total += t.w * t.h;
        s.tex_offsets.push_back(total);
        last_w = t.w;
    }
    s.tex_data.resize(s.tex_data.size() + 3 * (size_t)last_w, 0.f);        // zero padding: one row of the last texture
    if (s.n_tex == 0) s.tex_offsets.clear();   // the reference crashes without textures (D14); the new path just has none

    // material table: ids in first-seen order through the reference's map + find (scene.cpp:80-103, 192-203)
    std::map<Mat, unsigned, RefLess> ids;
    unsigned next = 0;
    for (const ObjMesh& m : obj.meshes) {
        const Mat mm = from_mtl(m.material);
        if (ids.count(mm) == 0) { ids.insert({mm, next}); next++; }
    }
    s.n_mats = (uint32_t)ids.size();
    std::vector<Mat> table(s.n_mats);
    for (const auto& kv : ids)
        if (kv.second < s.n_mats) table[kv.second] = kv.first;
    const std::vector<Mat> lookup = table;

    size_t n_tris = 0;
    for (const ObjMesh& m : obj.meshes) n_tris += m.vertices.size() / 3;
    s.tris.resize(n_tris);                                  // value-initialised: every byte zero, padding included
    if (obj.all_vertices.size() != 3 * n_tris) die("internal: vertex soup and triangle count disagree");
    std::vector<float> positions(9 * n_tris);
    uint32_t tri_base = 0;
    for (const ObjMesh& m : obj.meshes) {
        const Mat mm = from_mtl(m.material);
        const auto it = std::find_if(lookup.begin(), lookup.end(), [&](const Mat& x) { return ref_equal(x, mm); });
        if (it == lookup.end()) die("material of mesh `" + m.name + "` is not in the material table");
        const uint32_t mat_id = (uint32_t)(it - lookup.begin());
        if (!m.material.map_Kd.empty()) {
            const auto tf = std::find(tex_files.begin(), tex_files.end(), m.material.map_Kd);
            if (tf != tex_files.end()) table[mat_id].tex = (int)(tf - tex_files.begin());
        }
        s.mesh_report.emplace_back(m.name, m.vertices.size());
        const uint32_t n_mesh = (uint32_t)(m.vertices.size() / 3);
        const ObjVertex* const mv = m.vertices.data();
        crt::RefTriangle* const out_tris = s.tris.data() + tri_base;
        crt::parallel_for(n_mesh, [=](uint32_t lo, uint32_t hi) {       // scene.cpp:106-134, every triangle on its own
            for (uint32_t k = lo; k < hi; k++) {
                const ObjVertex &a = mv[3 * (size_t)k], &b = mv[3 * (size_t)k + 1], &c = mv[3 * (size_t)k + 2];
                crt::RefTriangle& t = out_tris[k];
                t.v1 = {a.p[0], a.p[1], a.p[2]};
                t.n1 = {a.n[0], a.n[1], a.n[2]}; t.n2 = {b.n[0], b.n[1], b.n[2]}; t.n3 = {c.n[0], c.n[1], c.n[2]};
                t.uv1 = {a.t[0], a.t[1]}; t.uv2 = {b.t[0], b.t[1]}; t.uv3 = {c.t[0], c.t[1]};
                t.e1 = {b.p[0] - a.p[0], b.p[1] - a.p[1], b.p[2] - a.p[2]};
                t.e2 = {c.p[0] - a.p[0], c.p[1] - a.p[1], c.p[2] - a.p[2]};
                t.materialId = mat_id;
                t.area = len3(cross3(t.e1, t.e2)) / 2.0f;
                t.triId = tri_base + k;
            }
        });
        tri_base += n_mesh;
    }
    {
        const ObjVertex* const av = obj.all_vertices.data();
        float* const pp = positions.data();
        crt::parallel_for((uint32_t)obj.all_vertices.size(), [=](uint32_t lo, uint32_t hi) {
            for (uint32_t i = lo; i < hi; i++) std::memcpy(pp + 3 * (size_t)i, av[i].p, 12);
        });
    }

    s.mat_floats.resize(12 * (size_t)s.n_mats);
    s.mat_indices.resize(2 * (size_t)s.n_mats);
    for (uint32_t m = 0; m < s.n_mats; m++) {              // sequential_renderer.cu:48-64
        float* f = &s.mat_floats[12 * m];
        std::memcpy(f, table[m].kd, 12); std::memcpy(f + 3, table[m].ka, 12); std::memcpy(f + 6, table[m].ks, 12);
        f[9] = table[m].ns; f[10] = table[m].ni; f[11] = table[m].diffuse;
        s.mat_indices[2 * m] = table[m].bsdf;
        s.mat_indices[2 * m + 1] = table[m].tex;
    }
    // BVH + light list: the product's host scene build, through the C ABI
    s.nodes.resize(std::max<size_t>(2 * n_tris, 1));
    uint32_t n_nodes = 0;
    if (crt_build_bvh(s.tris.data(), positions.data(), (uint32_t)n_tris, s.nodes.data(), &n_nodes) != CRT_OK) die(crt_last_error());
    s.nodes.resize(n_nodes);
    std::cout << "Total BVH Nodes: " << n_nodes << std::endl;
    s.lights.resize(std::max<size_t>(n_tris, 1));
    uint32_t n_lights = 0;
    if (crt_build_lights(s.tris.data(), (uint32_t)n_tris, s.mat_floats.data(), s.mat_indices.data(), s.n_mats, s.lights.data(),
                         &n_lights, &s.lights_area) != CRT_OK)
        die(crt_last_error());
    s.lights.resize(n_lights);
    return s;
}

void save_crts(const std::string& path, const SceneArrays& s, unsigned width, unsigned height) {
    FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) die("cannot write " + path);
    uint32_t hdr[16] = {0};
    std::memcpy(&hdr[0], "CRTS", 4);
    hdr[1] = 1; hdr[2] = (uint32_t)s.tris.size(); hdr[3] = (uint32_t)s.nodes.size(); hdr[4] = s.n_mats;
    hdr[5] = (uint32_t)s.lights.size(); hdr[6] = s.n_tex; hdr[7] = (uint32_t)(s.tex_data.size() / 3); hdr[8] = width; hdr[9] = height;
    std::memcpy(&hdr[10], &s.lights_area, 4);
    auto put = [&](const void* p, size_t n) { if (n && std::fwrite(p, 1, n, f) != n) die("write failed: " + path); };
    put(hdr, sizeof hdr);
    put(&s.camera, sizeof s.camera);
    put(s.tris.data(), s.tris.size() * sizeof(crt::RefTriangle));
    put(s.nodes.data(), s.nodes.size() * sizeof(crt::RefNode));
    put(s.mat_floats.data(), s.mat_floats.size() * 4);
    put(s.mat_indices.data(), s.mat_indices.size() * 4);
    put(s.lights.data(), s.lights.size() * 4);
    put(s.tex_dims.data(), s.tex_dims.size() * 4);
    if (s.n_tex) put(s.tex_offsets.data(), s.tex_offsets.size() * 4);
    else { const uint32_t zero = 0; put(&zero, 4); }
    put(s.tex_data.data(), s.tex_data.size() * 4);
    std::fclose(f);
}

}  // namespace crtcli
