// TEST INFRASTRUCTURE: spdlog stand-in. The reference only uses the two logging macros
// (83 SPDLOG_INFO / 23 SPDLOG_ERROR sites); neither affects results.
#pragma once
#include <cstdio>
#define SPDLOG_INFO(...) ((void)0)
#define SPDLOG_ERROR(...) (std::fprintf(stderr, "[sdflib-ref error] %s:%d\n", __FILE__, __LINE__))
