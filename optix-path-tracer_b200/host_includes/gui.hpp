// gui.hpp — the reference's ImGui/GLFW window (host_includes/gui.hpp, main.cpp:104-361)
// is out of scope: interactive Windows GUI, no GL on the box.  The headless CLI in
// ../app/main.cpp exposes the same App_State knobs as flags.
#pragma once
#include "scenes.hpp"
