// tests/stubs/utils/event.h -- TEST INFRASTRUCTURE: the event / handler pair of core/src/utils/event.h as far as the radio
// module's demodulator wrappers use it (bind, unbind, emit)
#pragma once
#include <algorithm>
#include <vector>

template <class T>
struct EventHandler {
    EventHandler() {}
    EventHandler(void (*h)(T, void*), void* c) : handler(h), ctx(c) {}
    void (*handler)(T, void*) = nullptr;
    void* ctx = nullptr;
};
template <class T>
class Event {
public:
    void emit(T value) { for (auto* h : handlers) { if (h->handler) { h->handler(value, h->ctx); } } }
    void bindHandler(EventHandler<T>* h) { handlers.push_back(h); }
    void unbindHandler(EventHandler<T>* h) { handlers.erase(std::remove(handlers.begin(), handlers.end(), h), handlers.end()); }
private:
    std::vector<EventHandler<T>*> handlers;
};
