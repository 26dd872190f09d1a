// Link stubs for the oracle build of the reference (TEST INFRASTRUCTURE — never shipped, never on the
// product path). The reference's game.cpp / run_game.cpp reference graphics::OpenGL and tree.cpp references
// GSL; neither is on the network hot path. OpenGL entry points are never reached when the Game has no
// graphics object; Dirichlet noise is only reached if MCTS is exercised.
#include <map>
#include <random>
#include <string>
#include <GL/glew.h>
#include <GLFW/glfw3.h>
#include <glm/glm.hpp>
#include "graphics/opengl.h"
#include <gsl/gsl_rng.h>
#include <gsl/gsl_randist.h>

std::map<GLFWwindow *, graphics::OpenGL *> graphics::OpenGL::_opengl_map;
graphics::OpenGL::~OpenGL() {}
void graphics::OpenGL::run() {}
void graphics::OpenGL::updateGraphics(game::Board *) {}
graphics::OpenGL *graphics::OpenGL::get_instance(game::Game *, const std::string &) { return nullptr; }
void graphics::OpenGL::run_graphics(game::Game *, const std::string &) {}

struct gsl_rng_type { int unused; };
struct gsl_rng { std::mt19937_64 engine; };
static const gsl_rng_type k_stub_type{0};
const gsl_rng_type *gsl_rng_mt19937 = &k_stub_type;
gsl_rng *gsl_rng_alloc(const gsl_rng_type *) { return new gsl_rng(); }
void gsl_rng_set(gsl_rng *r, unsigned long seed) { r->engine.seed(seed); }
void gsl_rng_free(gsl_rng *r) { delete r; }
void gsl_ran_dirichlet(const gsl_rng *r, size_t K, const double alpha[], double theta[]) {
  auto &eng = const_cast<gsl_rng *>(r)->engine;
  double total = 0.0;
  for (size_t i = 0; i < K; ++i) {
    std::gamma_distribution<double> g(alpha[i], 1.0);
    theta[i] = g(eng);
    total += theta[i];
  }
  for (size_t i = 0; i < K; ++i) theta[i] /= total;
}
