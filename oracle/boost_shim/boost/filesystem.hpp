// TEST INFRASTRUCTURE (oracle/_ref build only). boost::filesystem::path is used for extension()/stem()/string()
// in cpp/src/passband.cpp:10-14 only.
#pragma once
#include <filesystem>
namespace boost { namespace filesystem { using path = std::filesystem::path; } }
