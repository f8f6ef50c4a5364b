#ifndef MOCK_CAML_BIGARRAY_H
#define MOCK_CAML_BIGARRAY_H
#include "custom.h"
struct caml_ba_array {
  void *data;
  intnat num_dims;
  intnat flags;
  void *proxy;
  intnat dim[1];
};
#define Caml_ba_array_val(v) ((struct caml_ba_array *)Data_custom_val(v))
#define Caml_ba_data_val(v) (Caml_ba_array_val(v)->data)
/* a 1-D C-layout bigarray over caller-owned (malloc'ed) data, as Bigarray.Array1.create gives */
static inline value mock_caml_ba_alloc_1d(void *data, intnat n) {
  value v = mock_caml_alloc(1 + (sizeof(struct caml_ba_array) + sizeof(value) - 1) / sizeof(value), Custom_tag);
  Field(v, 0) = (value)0;
  struct caml_ba_array *b = Caml_ba_array_val(v);
  b->data = data, b->num_dims = 1, b->flags = 0, b->proxy = NULL, b->dim[0] = n;
  return v;
}
#endif
