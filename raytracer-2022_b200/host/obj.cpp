/*
 * obj.cpp -- Wavefront OBJ loading for the call the reference makes at scene.rs:368-375:
 *     tobj::load_obj(path, &LoadOptions { single_index: true, triangulate: true, .. })
 * tobj 3.2.2 is not vendored under /root/reference (Cargo.lock:484); this restates the part
 * of its behaviour the hot path depends on:
 *   - `v x y z` parsed as f32 (correctly rounded, like Rust's str::parse::<f32>);
 *   - `f` records with v, v/vt, v//vn or v/vt/vn corners, 1-based or negative indices;
 *     polygons are fan-triangulated (0, i, i+1) as tobj does;
 *   - a new model starts at each `o` / `g` statement; models without faces are dropped;
 *   - `single_index` re-indexes vertices per unique (v,vt,vn) triple but leaves triangle
 *     order and corner positions unchanged, so positions are emitted here as the position
 *     list referenced by the original `v` indices (geometry identical);
 *   - a missing `mtllib` file only makes the materials result an Err, which the reference
 *     ignores (scene.rs:379).
 */
#include "rt_host.h"

#include <cerrno>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>

namespace rt {

static bool parse_index(const char *tok, long n_pos, uint32_t &out) {
    // corner = v[/vt[/vn]]; only v matters here
    char *end = nullptr;
    long v = std::strtol(tok, &end, 10);
    if (end == tok) return false;
    long idx = v > 0 ? v - 1 : n_pos + v; // negative = relative to the end
    if (idx < 0 || idx >= n_pos) return false;
    out = (uint32_t)idx;
    return true;
}

std::vector<ObjMesh> load_obj(const std::string &path) {
    std::ifstream in(path);
    if (!in) throw Error{"load_obj: cannot open " + path};
    std::vector<float> positions; // global list, as in the file
    std::vector<ObjMesh> models;
    ObjMesh cur;
    cur.name = "unnamed_object";
    std::vector<uint32_t> cur_faces; // global position indices
    auto flush = [&]() {
        if (!cur_faces.empty()) {
            // compact the positions used by this model, keeping first-use order
            std::vector<int64_t> remap(positions.size() / 3, -1);
            for (uint32_t gi : cur_faces) {
                if (remap[gi] < 0) {
                    remap[gi] = (int64_t)(cur.positions.size() / 3);
                    cur.positions.push_back(positions[3 * gi]);
                    cur.positions.push_back(positions[3 * gi + 1]);
                    cur.positions.push_back(positions[3 * gi + 2]);
                }
                cur.indices.push_back((uint32_t)remap[gi]);
            }
            models.push_back(cur);
        }
        cur = ObjMesh();
        cur_faces.clear();
    };
    std::string line;
    size_t lineno = 0;
    while (std::getline(in, line)) {
        ++lineno;
        if (!line.empty() && line.back() == '\r') line.pop_back();
        std::istringstream ls(line);
        std::string tag;
        if (!(ls >> tag)) continue;
        if (tag == "v") {
            std::string xs[3];
            if (!(ls >> xs[0] >> xs[1] >> xs[2])) throw Error{"load_obj: bad vertex at line " + std::to_string(lineno)};
            for (int c = 0; c < 3; ++c) {
                errno = 0;
                char *end = nullptr;
                float f = std::strtof(xs[c].c_str(), &end); // correctly rounded in glibc
                if (end == xs[c].c_str()) throw Error{"load_obj: bad float at line " + std::to_string(lineno)};
                positions.push_back(f);
            }
        } else if (tag == "f") {
            std::vector<uint32_t> corners;
            std::string tok;
            while (ls >> tok) {
                uint32_t idx;
                if (!parse_index(tok.c_str(), (long)(positions.size() / 3), idx))
                    throw Error{"load_obj: bad face index at line " + std::to_string(lineno)};
                corners.push_back(idx);
            }
            if (corners.size() < 3) continue; // points / lines are not triangles
            for (size_t i = 1; i + 1 < corners.size(); ++i) {
                cur_faces.push_back(corners[0]);
                cur_faces.push_back(corners[i]);
                cur_faces.push_back(corners[i + 1]);
            }
        } else if (tag == "o" || tag == "g") {
            flush();
            std::string name;
            std::getline(ls, name);
            size_t b = name.find_first_not_of(" \t");
            cur.name = b == std::string::npos ? "unnamed_object" : name.substr(b);
        }
        // vt, vn, usemtl, mtllib, s, comments: not used by the hot path
    }
    flush();
    if (models.empty()) throw Error{"load_obj: no faces in " + path};
    return models;
}

} // namespace rt
