// unv_reader.cpp -- host-only ingest of the Salome-UNV subset the reference parses
// (src/mesh/MeshReaderSalomeUnv.cpp:267-448): blocks 2411 (nodes), 2412
// (fe_id 11 boundary edges / 41 triangles), 2467 (named groups).
//
// SURVEY 8(f) row 3 ("mesh ingest at scale"): the reference splits the file into std::string lists
// that it passes BY VALUE (parse_block(string_list sl), :428) and resolves group members and edges
// with linear searches (find_edge :14-22, Grid::findEdge grid.cpp:38-48), which makes multi-million
// cell inputs impractical.  Here the file is read once, tokenised in place with strtol/strtod (the
// same correctly rounded conversion as the reference's sscanf("%lf")), and element labels resolve
// through one flat table: O(file size).  The geometry (edges, normals, Gauss points, areas:
// :119-255) is built from these arrays by cfd-2d_b200/mesh.py exactly as before.
#include "../../include/cfd2d_fvm.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

extern void cfd2d_set_create_error(const std::string& s);   // fvm_api.cu

struct cfd2d_unv {
    std::vector<double> xy;            // nn * 2
    std::vector<int32_t> tris;         // nc * 3
    std::vector<int32_t> bedges;       // nbe * 2 (fe_id 11 elements, file order)
    struct Group { std::string name; std::vector<int64_t> cells; std::vector<int32_t> edges; };
    std::vector<Group> groups;         // sorted by name (std::map iteration order of the reference)
};

namespace {

struct Lines {
    std::vector<std::pair<const char*, const char*>> ln;   // [begin, end) without the newline / CR
    explicit Lines(std::vector<char>& buf) {
        const char* p = buf.data();
        const char* e = p + buf.size();
        while (p < e) {
            const char* q = (const char*)memchr(p, '\n', (size_t)(e - p));
            const char* le = q ? q : e;
            const char* t = le;
            if (t > p && t[-1] == '\r') --t;
            ln.emplace_back(p, t);
            if (!q) break;
            p = q + 1;
        }
    }
};

inline bool blank(const char* b, const char* e) {
    for (; b < e; ++b) if (!(*b == ' ' || *b == '\t' || *b == '\r' || *b == '\f' || *b == '\v')) return false;
    return true;
}

// a block delimiter: the FIRST occurrence of "-1" sits in the last two characters (:277-283)
inline bool is_delim(const char* b, const char* e) {
    if (e - b < 2) return false;
    for (const char* p = b; p + 1 < e; ++p)
        if (p[0] == '-' && p[1] == '1') return p == e - 2;
    return false;
}

// up to `want` integers of a line; returns how many were read (the line is NUL/newline terminated
// in the buffer copy, so strtol cannot run past it)
inline int ints(const char* b, const char* e, long long* out, int want) {
    int n = 0;
    const char* p = b;
    while (n < want && p < e) {
        while (p < e && (*p == ' ' || *p == '\t')) ++p;
        if (p >= e) break;
        char* q = nullptr;
        long long v = strtoll(p, &q, 10);
        if (q == p || q > e) break;
        out[n++] = v;
        p = q;
    }
    return n;
}

}  // namespace

extern "C" {

static int unv_read_impl(const char* path, cfd2d_unv** out);

int cfd2d_unv_read(const char* path, cfd2d_unv** out) {
    // no C++ exception may cross the C ABI (std::bad_alloc / length_error on hostile input)
    try {
        return unv_read_impl(path, out);
    } catch (const std::exception& ex) {
        if (out) *out = nullptr;
        cfd2d_set_create_error(std::string("cfd2d_unv_read: ") + ex.what());
        return CFD2D_EINVAL;
    } catch (...) {
        if (out) *out = nullptr;
        cfd2d_set_create_error("cfd2d_unv_read: unknown exception");
        return CFD2D_EINVAL;
    }
}

static int unv_read_impl(const char* path, cfd2d_unv** out) {
    if (!path || !out) { cfd2d_set_create_error("null argument"); return CFD2D_EINVAL; }
    *out = nullptr;
    FILE* f = fopen(path, "rb");
    if (!f) { cfd2d_set_create_error(std::string("cannot open ") + path); return CFD2D_EINVAL; }
    fseek(f, 0, SEEK_END);
    long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    std::vector<char> buf((size_t)sz + 1);
    if (sz > 0 && fread(buf.data(), 1, (size_t)sz, f) != (size_t)sz) { fclose(f); cfd2d_set_create_error("short read"); return CFD2D_EINVAL; }
    fclose(f);
    buf[(size_t)sz] = '\n';                                  // sentinel: strto* always stops
    Lines L(buf);
    if (!L.ln.empty() && L.ln.back().first >= buf.data() + sz) L.ln.pop_back();

    std::unique_ptr<cfd2d_unv> uown(new cfd2d_unv());      // freed on every early return and on exceptions
    cfd2d_unv* u = uown.get();
    // element label (0-based) -> +(cell index + 1) | -(edge element index + 1) | 0 = unknown
    std::vector<int64_t> elem;
    struct RawGroup { std::string name; std::vector<int64_t> labels; };
    std::vector<RawGroup> raw;
    std::string err;
    const size_t nl = L.ln.size();
    size_t i = 0;
    bool inside = false;
    size_t bstart = 0;
    auto parse_block = [&](size_t b0, size_t b1) {          // lines [b0, b1) of one block, b0 = kind line
        if (b0 >= b1) return;
        long long kind = 0;
        if (ints(L.ln[b0].first, L.ln[b0].second, &kind, 1) != 1) return;
        size_t k = b0 + 1;
        auto next_nonblank = [&]() -> long { while (k < b1 && blank(L.ln[k].first, L.ln[k].second)) ++k; return k < b1 ? (long)k++ : -1; };
        if (kind == 2411) {
            for (;;) {
                long r = next_nonblank();
                if (r < 0) break;
                if (k >= b1) { err = "2411: node record without coordinates"; return; }
                // both coordinates must sit on THIS line (the reference's sscanf("%lf %lf") stays on it too,
                // MeshReaderSalomeUnv.cpp:283-290): strtod skips newlines, so bound it by the line end
                const char* p = L.ln[k].first; const char* pe = L.ln[k].second; ++k;
                char* q = nullptr;
                double x = strtod(p, &q);
                if (q == p || q > pe) { err = "2411: bad coordinate line"; return; }
                const char* q0 = q;
                double y = strtod(q0, &q);
                if (q == q0 || q > pe) { err = "2411: coordinate line with a single number"; return; }
                u->xy.push_back(x); u->xy.push_back(y);
            }
        } else if (kind == 2412) {
            for (;;) {
                long r = next_nonblank();
                if (r < 0) break;
                long long t[2];
                if (ints(L.ln[r].first, L.ln[r].second, t, 2) != 2) { err = "2412: bad element record"; return; }
                const long long label = t[0] - 1, fe = t[1];
                if (label < 0) { err = "2412: element label < 1"; return; }
                // labels index a flat table: bound them by what the file could possibly hold (one element
                // takes >= 2 lines) times a generous factor, so a corrupt label cannot allocate gigabytes
                if ((unsigned long long)label > 64ull * (unsigned long long)nl + 1024ull) { err = "2412: element label out of range"; return; }
                if ((size_t)label >= elem.size()) elem.resize((size_t)label + 1 + elem.size() / 2, 0);
                if (fe == 11) {
                    if (k + 1 >= b1) { err = "2412: truncated beam element"; return; }
                    ++k;                                            // beam record 2
                    long long n[2];
                    if (ints(L.ln[k].first, L.ln[k].second, n, 2) != 2) { err = "2412: bad beam nodes"; return; }
                    ++k;
                    elem[(size_t)label] = -(int64_t)(u->bedges.size() / 2 + 1);
                    u->bedges.push_back((int32_t)(n[0] - 1)); u->bedges.push_back((int32_t)(n[1] - 1));
                } else if (fe == 41) {
                    if (k >= b1) { err = "2412: truncated triangle"; return; }
                    long long n[3];
                    if (ints(L.ln[k].first, L.ln[k].second, n, 3) != 3) { err = "2412: bad triangle nodes"; return; }
                    ++k;
                    elem[(size_t)label] = (int64_t)(u->tris.size() / 3 + 1);
                    u->tris.push_back((int32_t)(n[0] - 1)); u->tris.push_back((int32_t)(n[1] - 1)); u->tris.push_back((int32_t)(n[2] - 1));
                } else {
                    char b[64];
                    snprintf(b, sizeof b, "Unknown element type '%lld'.", fe);     // Exception::TYPE_MESH_UNV_UNKNOWN_ELEMENT
                    err = b;
                    return;
                }
            }
        } else if (kind == 2467) {
            for (;;) {
                long r = next_nonblank();
                if (r < 0) break;
                long long t[8];
                if (ints(L.ln[r].first, L.ln[r].second, t, 8) != 8) { err = "2467: bad group record"; return; }
                const long long n = t[7];
                if (n < 0 || (unsigned long long)n > 2ull * (unsigned long long)nl + 2ull) { err = "2467: bad entity count"; return; }
                if (k >= b1) { err = "2467: group without a name"; return; }
                const char* p = L.ln[k].first; const char* e = L.ln[k].second; ++k;
                while (p < e && (*p == ' ' || *p == '\t')) ++p;
                const char* q = p;
                while (q < e && !(*q == ' ' || *q == '\t')) ++q;
                RawGroup g;
                g.name.assign(p, q);
                g.labels.reserve((size_t)n);
                for (long long j = 0; j < n / 2; j++) {
                    if (k >= b1) { err = "2467: truncated group"; return; }
                    long long w[8];
                    if (ints(L.ln[k].first, L.ln[k].second, w, 8) != 8) { err = "2467: bad entity line"; return; }
                    ++k;
                    g.labels.push_back(w[1] - 1); g.labels.push_back(w[5] - 1);
                }
                if (n % 2 == 1) {
                    if (k >= b1) { err = "2467: truncated group"; return; }
                    long long w[2];
                    if (ints(L.ln[k].first, L.ln[k].second, w, 2) != 2) { err = "2467: bad entity line"; return; }
                    ++k;
                    g.labels.push_back(w[1] - 1);
                }
                // a later group with the same name replaces the earlier one (dict / std::map assignment)
                bool replaced = false;
                for (auto& o : raw) if (o.name == g.name) { o.labels.swap(g.labels); replaced = true; break; }
                if (!replaced) raw.push_back(std::move(g));
            }
        }
    };
    for (; i < nl && err.empty(); i++) {
        if (is_delim(L.ln[i].first, L.ln[i].second)) {
            if (inside) { parse_block(bstart, i); inside = false; }
            else { inside = true; bstart = i + 1; }
        }
    }
    if (err.empty() && inside && bstart < nl) parse_block(bstart, nl);
    if (!err.empty()) { cfd2d_set_create_error(err); return CFD2D_EINVAL; }
    std::sort(raw.begin(), raw.end(), [](const RawGroup& a, const RawGroup& b) { return a.name < b.name; });
    for (auto& g : raw) {
        cfd2d_unv::Group o;
        o.name = g.name;
        for (long long l : g.labels) {
            if (l < 0 || (size_t)l >= elem.size()) continue;
            int64_t v = elem[(size_t)l];
            if (v > 0) o.cells.push_back(v - 1);
            else if (v < 0) { size_t e = (size_t)(-v - 1); o.edges.push_back(u->bedges[2 * e]); o.edges.push_back(u->bedges[2 * e + 1]); }
        }
        u->groups.push_back(std::move(o));
    }
    *out = uown.release();
    return CFD2D_OK;
}

void cfd2d_unv_counts(const cfd2d_unv* u, int64_t* nn, int64_t* nc, int64_t* nbe, int32_t* ngroups) {
    if (nn) *nn = u ? (int64_t)u->xy.size() / 2 : 0;
    if (nc) *nc = u ? (int64_t)u->tris.size() / 3 : 0;
    if (nbe) *nbe = u ? (int64_t)u->bedges.size() / 2 : 0;
    if (ngroups) *ngroups = u ? (int32_t)u->groups.size() : 0;
}

void cfd2d_unv_copy(const cfd2d_unv* u, double* xy, int32_t* tris, int32_t* bedges) {
    if (!u) return;
    if (xy && !u->xy.empty()) memcpy(xy, u->xy.data(), u->xy.size() * sizeof(double));
    if (tris && !u->tris.empty()) memcpy(tris, u->tris.data(), u->tris.size() * sizeof(int32_t));
    if (bedges && !u->bedges.empty()) memcpy(bedges, u->bedges.data(), u->bedges.size() * sizeof(int32_t));
}

const char* cfd2d_unv_group_name(const cfd2d_unv* u, int g) {
    return (u && g >= 0 && (size_t)g < u->groups.size()) ? u->groups[(size_t)g].name.c_str() : "";
}

void cfd2d_unv_group_counts(const cfd2d_unv* u, int g, int64_t* ncells, int64_t* nedges) {
    bool ok = u && g >= 0 && (size_t)g < u->groups.size();
    if (ncells) *ncells = ok ? (int64_t)u->groups[(size_t)g].cells.size() : 0;
    if (nedges) *nedges = ok ? (int64_t)u->groups[(size_t)g].edges.size() / 2 : 0;
}

void cfd2d_unv_group_copy(const cfd2d_unv* u, int g, int64_t* cells, int32_t* edge_nodes) {
    if (!(u && g >= 0 && (size_t)g < u->groups.size())) return;
    const auto& G = u->groups[(size_t)g];
    if (cells && !G.cells.empty()) memcpy(cells, G.cells.data(), G.cells.size() * sizeof(int64_t));
    if (edge_nodes && !G.edges.empty()) memcpy(edge_nodes, G.edges.data(), G.edges.size() * sizeof(int32_t));
}

void cfd2d_unv_free(cfd2d_unv* u) { delete u; }

}  // extern "C"
