/* Hand-written stand-in for the header the reference's CMake step
 * `generate_export_header` would emit (src/CMakeLists.txt:113 of the reference).
 * The oracle build links the reference statically into a private shared object,
 * so the export macros are empty.  TEST INFRASTRUCTURE ONLY (see oracle/README.md). */
#ifndef YAFARAY_C_API_EXPORT_H
#define YAFARAY_C_API_EXPORT_H
#define YAFARAY_C_API_EXPORT
#define YAFARAY_C_API_NO_EXPORT
#define YAFARAY_C_API_DEPRECATED
#endif
