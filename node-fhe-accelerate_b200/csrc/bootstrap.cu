// placeholder entry points (replaced by the real kernels)
#include "runtime.hpp"
using namespace fheb;
#define NI return set_error(FHEB_ERR_NATIVE, "not implemented yet")
extern "C" {
int fheb_boot_key_create(const fheb_ntt_plan*, const fheb_boot_params*, const uint64_t*, fheb_boot_key**) { NI; }
int fheb_boot_key_set_ksk(fheb_boot_key*, const uint64_t*, size_t, uint32_t, uint32_t, uint32_t) { NI; }
int fheb_boot_key_destroy(fheb_boot_key*) { return 0; }
int fheb_external_product_batch(const fheb_boot_key*, uint32_t, const uint64_t*, uint64_t*, size_t, void*) { NI; }
int fheb_cmux_batch(const fheb_boot_key*, uint32_t, const uint64_t*, const uint64_t*, uint64_t*, size_t, void*) { NI; }
int fheb_blind_rotate_batch(const fheb_boot_key*, const uint64_t*, const uint64_t*, uint64_t*, size_t, void*) { NI; }
int fheb_sample_extract_batch(const fheb_boot_key*, const uint64_t*, uint64_t*, size_t, void*) { NI; }
int fheb_key_switch_batch(const fheb_boot_key*, const uint64_t*, uint64_t*, size_t, void*) { NI; }
int fheb_bootstrap_batch(const fheb_boot_key*, const uint64_t*, const uint64_t*, uint64_t*, size_t, void*) { NI; }
int fheb_make_test_poly(const fheb_ntt_plan*, int, uint64_t, uint64_t, uint64_t*) { NI; }
}
