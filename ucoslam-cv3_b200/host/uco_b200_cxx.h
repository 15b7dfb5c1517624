// uco_b200_cxx.h — what every C++ adapter shares: a RAII context and the status -> exception translation.
// The reference reports errors by throwing std::runtime_error (src/featureextractors/feature2dserializable.cpp:71,
// 3rdparty/fbow/fbow/fbow.cpp:52-54); so do the adapters.
#pragma once
#include <memory>
#include <stdexcept>
#include <string>
#include "ucoslam_b200.h"

namespace uco_b200 {

inline void check(uco_b200_ctx* ctx, int rc) {
    if (rc != UCO_OK) throw std::runtime_error(std::string("ucoslam_b200: ") + uco_b200_last_error(ctx));
}

// one context (CUDA stream + workspaces) per owning object / calling thread; not thread safe by design
class Context {
public:
    explicit Context(int device = 0) : _ctx(uco_b200_create(device, 0)) {
        if (!_ctx) throw std::runtime_error("ucoslam_b200: no usable CUDA device (there is no CPU fallback)");
    }
    ~Context() { uco_b200_destroy(_ctx); }
    Context(const Context&) = delete;
    Context& operator=(const Context&) = delete;
    uco_b200_ctx* get() const { return _ctx; }
    void check(int rc) const { uco_b200::check(_ctx, rc); }

private:
    uco_b200_ctx* _ctx;
};

}  // namespace uco_b200
