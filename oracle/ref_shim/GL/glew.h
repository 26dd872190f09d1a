// Header stub so the reference's graphics/opengl.h parses without GLEW installed (oracle build only).
#pragma once
typedef unsigned int GLuint;
