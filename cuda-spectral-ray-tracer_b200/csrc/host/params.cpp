// CLI parameters: same flags, defaults and error behaviour as the reference's
// io/params.h:21-315 (class parameters + param_manager singleton).
#include "srt_host.hpp"
#include <iostream>
#include <sstream>

namespace srt {

static const char* const kSceneNames[] = {"Cornell Box", "Prism World", "Different Materials"};  // params.h:19
static std::unique_ptr<Params> g_params;
static thread_local std::string g_error;

void set_error(const std::string& msg) { g_error = msg; }
const std::string& last_error() { return g_error; }

Params& Params::instance() {
    if (!g_params) g_params.reset(new Params());
    return *g_params;
}
void Params::reset_instance() { g_params.reset(); }

void Params::reset_yres() {  // params.h:176-180: uint(xres / ar), at least 1
    yres = static_cast<unsigned>(xres / ar);
    if (yres < 1) yres = 1;
}
unsigned Params::get_xcsize() const {  // params.h:53-57
    unsigned v = xcsize == 0 ? ycsize : xcsize;
    return v == 0 ? xres : v;
}
unsigned Params::get_ycsize() const {  // params.h:59-63
    unsigned v = ycsize == 0 ? xcsize : ycsize;
    return v == 0 ? yres : v;
}
const std::string& Params::get_title() const {  // params.h:28-31
    title_cache = image_title.empty() ? std::string(kSceneNames[scene < 3 ? scene : 0]) : image_title;
    return title_cache;
}

static bool parse_uint(const char* what, const std::string& s, unsigned& out) {
    try {
        out = static_cast<unsigned>(std::stoul(s));
        return true;
    } catch (...) {
        std::cerr << "Error while parsing " << what << " arg value, keeping previous (default most likely) value" << std::endl;
        return false;
    }
}

static float parse_ar(const std::string& text) {  // params.h:182-195: "a" or "a/b"
    std::stringstream ss(text);
    std::string tok;
    std::getline(ss, tok, '/');
    float ar = std::stof(tok);
    if (std::getline(ss, tok, '/')) {
        ar /= std::stof(tok);
        if (std::getline(ss, tok, '/'))
            std::cout << "Characters inserted after aspect ratio's denominator will be ignored, computed AR value is: " << ar
                      << std::endl;
    }
    return ar;
}

void Params::parse(int argc, char** argv) {  // params.h:236-304
    for (int i = 1; i < argc; i++) {
        std::string arg(argv[i]);
        const bool has_value = i + 1 < argc;
        auto is = [&](const char* a, const char* b) { return has_value && (arg == a || arg == b); };
        if (is("-t", "--title")) image_title = argv[++i];
        else if (is("-lsub", "--log-subdir")) log_subdir = argv[++i];
        else if (is("-s", "--scene")) parse_uint("scene", argv[++i], scene);
        else if (is("-xr", "--xres")) { if (parse_uint("xres", argv[++i], xres)) reset_yres(); }
        else if (is("-ar", "--aspect-ratio")) {
            try {
                ar = parse_ar(argv[++i]);
                reset_yres();
            } catch (...) {
                std::cerr << "Error while parsing aspect-ratio arg value, keeping previous (default most likely) value" << std::endl;
            }
        }
        else if (is("-xc", "--xcsize")) parse_uint("xcsize", argv[++i], xcsize);
        else if (is("-yc", "--ycsize")) parse_uint("ycsize", argv[++i], ycsize);
        else if (is("-ns", "--nsamples")) parse_uint("n_samples", argv[++i], n_samples);
        else if (is("-bl", "--bounce-limit")) parse_uint("bounce-limit", argv[++i], bounce_limit);
        else if (arg == "--do-log") do_log = true;
        else if (arg == "--no-show") show_render = false;
        else if (arg == "--save") do_save = true;
        else std::cout << "Unkown argument name: " << arg << std::endl;
    }
}

}  // namespace srt
