#ifndef MOCK_CAML_CUSTOM_H
#define MOCK_CAML_CUSTOM_H
#include "mlvalues.h"
typedef uintptr_t uintnat_placeholder_unused;
struct custom_operations {
  const char *identifier;
  void (*finalize)(value);
  int (*compare)(value, value);
  intnat (*hash)(value);
  void (*serialize)(value, uintnat_placeholder_unused *, uintnat_placeholder_unused *);
  uintptr_t (*deserialize)(void *);
  int (*compare_ext)(value, value);
  const void *fixed_length;
};
#define custom_compare_default NULL
#define custom_hash_default NULL
#define custom_serialize_default NULL
#define custom_deserialize_default NULL
#define custom_compare_ext_default NULL
#define custom_fixed_length_default NULL
#define Data_custom_val(v) ((void *)&Field((v), 1))
static inline value caml_alloc_custom(struct custom_operations *ops, uintptr_t size, mlsize_t mem, mlsize_t max) {
  (void)mem, (void)max;
  value v = mock_caml_alloc(1 + (size + sizeof(value) - 1) / sizeof(value), Custom_tag);
  Field(v, 0) = (value)ops;
  return v;
}
static inline void mock_caml_finalize(value v) {  /* what the GC would do when the block dies */
  struct custom_operations *ops = (struct custom_operations *)Field(v, 0);
  if (ops && ops->finalize) ops->finalize(v);
}
#endif
