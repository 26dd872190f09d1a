// Header stub so the reference's graphics/opengl.h parses without GLFW installed (oracle build only).
#pragma once
struct GLFWwindow;
