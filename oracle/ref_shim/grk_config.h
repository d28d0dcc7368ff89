/* Hand-written stand-in for the header the reference's CMake would generate from
 * src/lib/core/grk_config.h.cmake.in -- only what the few translation units compiled by
 * oracle/Makefile need.  Test infrastructure only. */
#pragma once
#define GRK_VERSION_MAJOR 20
#define GRK_VERSION_MINOR 4
#define GRK_VERSION_BUILD 12
#define GRK_PACKAGE_VERSION "20.4.12"
