// tests/stubs/gui/widgets/waterfall.h -- TEST INFRASTRUCTURE.  The few ImGui names the reference's demodulator wrappers
// touch (widgets return "unchanged"), so that those headers compile without the GUI.
#pragma once
#include <string>

struct ImVec2 { float x = 0.0f, y = 0.0f; };
namespace ImGui {
    class WaterfallVFO {
    public:
        enum { REF_LOWER, REF_CENTER, REF_UPPER, _REF_COUNT };
    };
    inline bool Checkbox(const char*, bool*) { return false; }
    inline bool SliderFloat(const char*, float*, float, float) { return false; }
    inline void LeftLabel(const char*) {}
    inline void SetNextItemWidth(float) {}
    inline float GetCursorPosX() { return 0.0f; }
    inline ImVec2 GetContentRegionAvail() { return ImVec2(); }
}
