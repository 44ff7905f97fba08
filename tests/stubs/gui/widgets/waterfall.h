// tests/stubs/gui/widgets/waterfall.h -- TEST INFRASTRUCTURE.  The ImGui / waterfall names the reference's demodulator wrappers
// touch (widgets return "unchanged", drawing does nothing), so that those headers compile without the GUI.
#pragma once
#include <string>
#include <vector>
#include <utils/event.h>

struct ImVec2 {
    float x = 0.0f, y = 0.0f;
    ImVec2() {}
    ImVec2(float x_, float y_) : x(x_), y(y_) {}
};
typedef unsigned int ImU32;
#define IM_COL32(r, g, b, a) ((ImU32)(((ImU32)(a) << 24) | ((ImU32)(b) << 16) | ((ImU32)(g) << 8) | (ImU32)(r)))
struct ImDrawList {
    int rects = 0, texts = 0;
    std::string last_text;
    void AddRectFilled(const ImVec2&, const ImVec2&, ImU32) { rects++; }
    void AddText(const ImVec2&, ImU32, const char* t) { texts++; last_text = t; }
};
struct ImGuiWindow { ImDrawList list; ImDrawList* DrawList = &list; };
enum { ImGuiTableFlags_SizingFixedFit = 1, ImGuiTableFlags_RowBg = 2, ImGuiTableFlags_Borders = 4 };
namespace style { static float uiScale = 1.0f; }
namespace ImGui {
    class WaterfallVFO {
    public:
        enum { REF_LOWER, REF_CENTER, REF_UPPER, _REF_COUNT };
    };
    class WaterFall {
    public:
        struct FFTRedrawArgs {
            ImVec2 min, max;
            double lowFreq = 0, highFreq = 0, freqToPixelRatio = 0, pixelToFreqRatio = 0;
            ImGuiWindow* window = nullptr;
        };
        Event<FFTRedrawArgs> onFFTRedraw;
    };
    inline bool Checkbox(const char*, bool*) { return false; }
    inline bool SliderFloat(const char*, float*, float, float) { return false; }
    inline bool Combo(const char*, int*, const char*) { return false; }
    inline void LeftLabel(const char*) {}
    inline void SameLine() {}
    inline void FillWidth() {}
    inline void SetNextItemWidth(float) {}
    inline float GetCursorPosX() { return 0.0f; }
    inline ImVec2 GetContentRegionAvail() { return ImVec2(); }
    inline ImVec2 CalcTextSize(const char* t) { return ImVec2(7.0f * (float)std::string(t).size(), 13.0f); }
    inline bool BeginTable(const char*, int, int = 0) { return true; }
    inline void EndTable() {}
    inline void TableNextRow() {}
    inline void TableSetColumnIndex(int) {}
    inline void TextUnformatted(const char*) {}
    inline void Text(const char*, ...) {}
}
namespace gui { static ImGui::WaterFall waterfall; }
