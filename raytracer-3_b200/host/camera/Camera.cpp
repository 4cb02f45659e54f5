/* camera/Camera.cpp — see Camera.hpp. */
#include <cmath>

#include "Camera.hpp"

using namespace RayTracer;

Camera::Camera() : frame(nullptr) {}

Camera::Camera(const Camera& other)
    : origin(other.origin), horizontal(other.horizontal), vertical(other.vertical), lower_left_corner(other.lower_left_corner),
      lens_radius(other.lens_radius), lens_u(other.lens_u), lens_v(other.lens_v), frame(other.frame ? new Frame(*other.frame) : nullptr) {}

Camera::~Camera() { delete this->frame; }

void Camera::update(uint32_t width, uint32_t height, float focal_length, float viewport_width, float viewport_height) {
    delete this->frame;
    this->frame = new Frame(width, height);
    this->origin = glm::vec3(0.0f, 0.0f, 0.0f);
    this->horizontal = glm::vec3(viewport_width, 0.0f, 0.0f);
    this->vertical = glm::vec3(0.0f, viewport_height, 0.0f);
    this->lower_left_corner = this->origin - this->horizontal / glm::vec3(2.0f) - this->vertical / glm::vec3(2.0f) - glm::vec3(0.0f, 0.0f, focal_length);
    this->lens_radius = 0.0f;
}

void Camera::look_at(uint32_t width, uint32_t height, const glm::vec3& from, const glm::vec3& at, const glm::vec3& up, float vfov_degrees,
                     float aperture, float focus_distance) {
    delete this->frame;
    this->frame = new Frame(width, height);
    const double theta = (double) vfov_degrees * (3.14159265358979323846 / 180.0);
    const double vh = 2.0 * std::tan(theta / 2.0), vw = ((double) width / (double) height) * vh;
    auto norm = [](double* v) { double l = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); v[0] /= l; v[1] /= l; v[2] /= l; };
    double w[3] = { (double) from.x - at.x, (double) from.y - at.y, (double) from.z - at.z };
    norm(w);
    double u[3] = { up.y * w[2] - up.z * w[1], up.z * w[0] - up.x * w[2], up.x * w[1] - up.y * w[0] };
    norm(u);
    double v[3] = { w[1] * u[2] - w[2] * u[1], w[2] * u[0] - w[0] * u[2], w[0] * u[1] - w[1] * u[0] };
    const double f[3] = { from.x, from.y, from.z };
    double hor[3], ver[3], llc[3];
    for (int i = 0; i < 3; i++) {
        hor[i] = focus_distance * vw * u[i];
        ver[i] = focus_distance * vh * v[i];
        llc[i] = f[i] - hor[i] / 2 - ver[i] / 2 - focus_distance * w[i];
    }
    this->origin = from;
    this->horizontal = glm::vec3(hor[0], hor[1], hor[2]);
    this->vertical = glm::vec3(ver[0], ver[1], ver[2]);
    this->lower_left_corner = glm::vec3(llc[0], llc[1], llc[2]);
    this->lens_radius = aperture / 2.0f;
    this->lens_u = glm::vec3(u[0], u[1], u[2]);
    this->lens_v = glm::vec3(v[0], v[1], v[2]);
}
