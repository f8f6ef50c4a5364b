#ifndef MOCK_CAML_ALLOC_H
#define MOCK_CAML_ALLOC_H
#include "mlvalues.h"
static inline value caml_alloc(mlsize_t wosize, int tag) { return mock_caml_alloc(wosize, tag); }
static inline value caml_alloc_tuple(mlsize_t n) { return mock_caml_alloc(n, 0); }
#endif
