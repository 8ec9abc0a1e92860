/*
 * TEST INFRASTRUCTURE ONLY (oracle/_ref build). Not part of the product.
 *
 * Force-included (-include) in front of every reference translation unit.
 *
 * Why it exists: the reference's fhe_types.h (cpp/include/fhe_types.h:29-33,
 * 45-57) and parameter_set.h (cpp/include/parameter_set.h:24-28,70-166) both
 * define fhe_accelerate::SecurityLevel and fhe_accelerate::ParameterSet, and
 * ntt_processor.h:18 + key_manager.h:22 pull both into the same TU, which no
 * compiler accepts.  Because fhe_types.h is `#pragma once`, including it here
 * first with the two clashing names renamed makes the later include a no-op
 * and leaves parameter_set.h's definitions (the ones the engines use) as the
 * only ones visible.  fhe_types.h:19-26 likewise clashes with the
 * HardwareCapabilities of adaptive_dispatcher.h:97-108 (no file on the path
 * uses either), so it is renamed the same way.  No reference source is
 * modified or copied.
 */
#pragma once
#define SecurityLevel SecurityLevel__fhe_types_duplicate
#define ParameterSet ParameterSet__fhe_types_duplicate
#define HardwareCapabilities HardwareCapabilities__fhe_types_duplicate
#include "fhe_types.h"
#undef SecurityLevel
#undef ParameterSet
#undef HardwareCapabilities
