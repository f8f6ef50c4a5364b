#ifndef MOCK_CAML_FAIL_H
#define MOCK_CAML_FAIL_H
#include "mlvalues.h"
static inline void caml_failwith(const char *msg) { mock_caml_raise("Failure", msg); }
static inline void caml_invalid_argument(const char *msg) { mock_caml_raise("Invalid_argument", msg); }
#endif
