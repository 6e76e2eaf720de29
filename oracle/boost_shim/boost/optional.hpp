// TEST INFRASTRUCTURE (oracle/_ref build only). Minimal stand-in for boost::optional: the reference
// only default-constructs, tests with operator!, assigns and dereferences (cpp/include/freddi_state.hpp:261-264,365-376).
#pragma once
#include <optional>
namespace boost {
template <class T> using optional = std::optional<T>;
}
