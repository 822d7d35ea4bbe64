// Flag parsing of the reference CLI (main.cpp:139-214): same flags, same defaults, numeric flags parsed as
// float and widened to double exactly like args::ValueFlag<float> does.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "pcd.h"
#include "pcd_host.h"

namespace pcdh {
std::string g_host_err;
}

static bool parse_float(const char *s, double &out) {
    char *end = nullptr;
    const float f = strtof(s, &end);
    if (end == s || *end != '\0') return false;
    out = (double)f;
    return true;
}

static bool parse_int(const char *s, int &out) {
    char *end = nullptr;
    const long v = strtol(s, &end, 10);
    if (end == s || *end != '\0') return false;
    out = (int)v;
    return true;
}

extern "C" int pcd_host_parse_cli(int argc, const char *const *argv, pcd_cli_options *o) {
    memset(o, 0, sizeof(*o));
    strcpy(o->progress_out, "./");   // main.cpp:176
    strcpy(o->output, "./");         // :177
    o->res_w = 100;                  // :178
    o->mesh_width = 1.0;             // :179
    o->focal_l = 1.5;                // :180
    o->thickness = 0.2;              // :181
    o->threads = 1;                  // :182
    o->conv_tres = 0.01;             // :183
    o->device = 0;
    o->solver_path = PCD_SOLVER_AUTO;
    o->gpus = 1;
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        if (a == "-h" || a == "--help") { o->help = 1; continue; }
        if (a.rfind("--", 0) != 0) { pcdh::g_host_err = "Flag could not be matched: " + a; return 1; }
        std::string name = a.substr(2), value;
        const size_t eq = name.find('=');
        bool have = false;
        if (eq != std::string::npos) { value = name.substr(eq + 1); name = name.substr(0, eq); have = true; }
        if (name == "quiet") { o->quiet = 1; continue; }
        if (!have) {
            if (i + 1 >= argc) { pcdh::g_host_err = "Flag '" + name + "' requires an argument but received none"; return 1; }
            value = argv[++i];
        }
        bool ok = true;
        auto str = [&](char *dst) { snprintf(dst, 1024, "%s", value.c_str()); };
        if (name == "input_png") str(o->input_png);
        else if (name == "progress_out") { str(o->progress_out); o->has_progress_out = 1; }
        else if (name == "output") str(o->output);
        else if (name == "res_w") ok = parse_int(value.c_str(), o->res_w);
        else if (name == "mesh_width") ok = parse_float(value.c_str(), o->mesh_width);
        else if (name == "focal_l") ok = parse_float(value.c_str(), o->focal_l);
        else if (name == "thickness") ok = parse_float(value.c_str(), o->thickness);
        else if (name == "threads") ok = parse_int(value.c_str(), o->threads);
        else if (name == "conv_tres") ok = parse_float(value.c_str(), o->conv_tres);
        else if (name == "device") ok = parse_int(value.c_str(), o->device);
        else if (name == "gpus") ok = parse_int(value.c_str(), o->gpus) && o->gpus >= 1;
        else if (name == "solver_path") {
            if (value == "auto") o->solver_path = PCD_SOLVER_AUTO;
            else if (value == "streaming") o->solver_path = PCD_SOLVER_STREAMING;
            else if (value == "resident") o->solver_path = PCD_SOLVER_RESIDENT;
            else if (value == "tiled") o->solver_path = PCD_SOLVER_TILED;
            else if (value == "dct") o->solver_path = PCD_SOLVER_DCT;   // opt-in direct backend
            else ok = false;
        } else { pcdh::g_host_err = "Flag could not be matched: " + name; return 1; }
        if (!ok) { pcdh::g_host_err = "Argument '" + name + "' received invalid value type '" + value + "'"; return 1; }
    }
    return 0;
}
