// tests/stubs/gui/widgets/symbol_diagram.h -- TEST INFRASTRUCTURE: the buffer contract of ImGui::SymbolDiagram
// (core/src/gui/widgets/symbol_diagram.h) without the drawing; `frames` counts released buffers
#pragma once
#include <mutex>
#include <vector>
#include <gui/widgets/waterfall.h>

namespace ImGui {
    class SymbolDiagram {
    public:
        SymbolDiagram(float scale = 1.0f, int count = 1024) : buffer((size_t)count, 0.0f), _scale(scale) {}
        void draw(const ImVec2& = ImVec2(0, 0)) {}
        float* acquireBuffer() { mtx.lock(); return buffer.data(); }
        void releaseBuffer() { frames++; mtx.unlock(); }
        std::vector<float> lines;
        int frames = 0;
    private:
        std::mutex mtx;
        std::vector<float> buffer;
        float _scale;
    };
}
