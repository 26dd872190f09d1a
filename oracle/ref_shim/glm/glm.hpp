// Header stub so the reference's graphics/opengl.h parses without GLM installed (oracle build only).
#pragma once
namespace glm { struct mat4 { float m[16]; }; }
