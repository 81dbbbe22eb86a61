// oracle/shadow/ScopePanel.h -- TEST INFRASTRUCTURE ONLY.
//
// Stands in for the reference's OpenGL panel header (src/panel/ScopePanel.h, which pulls in wx) on the include path when
// oracle/build_ref.py compiles src/process/ScopeVisualProcessor.cpp UNMODIFIED.  The processor uses one thing of the panel:
// the ScopeMode enumeration it stamps on its render data (ScopePanel.h:11).
#pragma once
// (the real header brings these in transitively)
#include <cmath>
#include <cstdlib>
class ScopePanel {
public:
    typedef enum ScopeMode { SCOPE_MODE_Y, SCOPE_MODE_2Y, SCOPE_MODE_XY } ScopeMode;
};
