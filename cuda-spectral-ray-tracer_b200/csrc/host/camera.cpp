// Pinhole / thin-lens camera set-up with the float operation order of the reference's
// rendering/camera.cu:7-58 (host code there too), producing the camera_data fields the
// render kernels consume (rendering/rendering.cuh:19-37).
#include "srt_host.hpp"
#include <cmath>

namespace srt {
namespace {
constexpr float kPi = 3.1415926535897932385f;  // utils/utility.h:12
inline vec3f add(vec3f a, vec3f b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline vec3f sub(vec3f a, vec3f b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline vec3f mul(float t, vec3f v) { return {t * v.x, t * v.y, t * v.z}; }
inline vec3f divs(vec3f v, float t) { return mul(1 / t, v); }  // vec3 operator/ = reciprocal multiply (math/vec3.cuh:144-147)
inline vec3f cross(vec3f u, vec3f v) { return {u.y * v.z - u.z * v.y, u.z * v.x - u.x * v.z, u.x * v.y - u.y * v.x}; }
inline float len(vec3f v) { return std::sqrt(v.x * v.x + v.y * v.y + v.z * v.z); }
inline vec3f unit(vec3f v) { return divs(v, len(v)); }
inline srt_vec3 out(vec3f v) { return {v.x, v.y, v.z}; }
inline float radians(float deg) { return deg * kPi / 180.0f; }
}  // namespace

srt_camera CameraBuilder::build(uint32_t w, uint32_t h) const {
    srt_camera c{};
    c.width = w;
    c.height = h;
    const vec3f center = lookfrom;
    const float theta = radians(vfov);
    const float half_h = std::tan(theta / 2.0f) * focus_dist;
    const float viewport_h = 2.0f * half_h;
    const float viewport_w = viewport_h * ((float)w / (float)h);
    const vec3f bw = unit(sub(lookfrom, lookat));
    const vec3f bu = unit(cross(vup, bw));
    const vec3f bv = cross(bw, bu);
    const vec3f viewport_u = mul(viewport_w, bu);
    const vec3f viewport_v = mul(viewport_h, vec3f(-bv.x, -bv.y, -bv.z));
    const vec3f du = divs(viewport_u, (float)(int)w);
    const vec3f dv = divs(viewport_v, (float)(int)h);
    const vec3f upper_left = sub(sub(sub(center, mul(focus_dist, bw)), divs(viewport_u, 2)), divs(viewport_v, 2));
    const vec3f p00 = add(upper_left, mul(0.5f, add(du, dv)));
    const float defocus_radius = focus_dist * std::tan(radians(defocus_angle / 2));
    c.pixel_delta_u = out(du);
    c.pixel_delta_v = out(dv);
    c.pixel00_loc = out(p00);
    c.defocus_angle = defocus_angle;
    c.camera_center = out(center);
    c.defocus_disk_u = out(mul(defocus_radius, bu));
    c.defocus_disk_v = out(mul(defocus_radius, bv));
    c.background = out(background);
    return c;
}

}  // namespace srt
