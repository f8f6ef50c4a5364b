/* tests/mock_caml — a stand-in for the OCaml 5 C runtime headers, just enough of <caml/...> to COMPILE, LINK and
 * RUN integration/ocaml/ptb_stubs.c without an OCaml toolchain (this image has none).  Test infrastructure only.
 * Values: immediates are (n << 1) | 1; blocks are pointers to their first field with a header word in front
 * ((wosize << 10) | tag) — the real runtime's representation, minus the GC (nothing is ever moved or freed).
 * caml_failwith / caml_invalid_argument longjmp to a handler the test driver installs. */
#ifndef MOCK_CAML_MLVALUES_H
#define MOCK_CAML_MLVALUES_H
#include <setjmp.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef intptr_t value;
typedef uintptr_t mlsize_t;
typedef intptr_t intnat;
typedef uintptr_t header_t;

#define CAMLprim
#define Val_int(x) ((value)(((intnat)(x) << 1) + 1))
#define Int_val(v) ((int)((v) >> 1))
#define Val_long(x) ((value)(((intnat)(x) << 1) + 1))
#define Long_val(v) ((intnat)((v) >> 1))
#define Store_field(b, i, v) (Field(b, i) = (v))
#define Val_unit Val_int(0)
#define Is_block(v) (((v)&1) == 0)
#define Hd_val(v) (((header_t *)(v))[-1])
#define Wosize_val(v) ((mlsize_t)(Hd_val(v) >> 10))
#define Tag_val(v) ((int)(Hd_val(v) & 0xFF))
#define Field(v, i) (((value *)(v))[i])
#define Double_tag 253
#define Custom_tag 255
#define Double_val(v) (*(double *)(v))

static inline value mock_caml_alloc(mlsize_t wosize, int tag) {
  header_t *b = (header_t *)calloc(wosize + 1, sizeof(value));
  b[0] = ((header_t)wosize << 10) | (header_t)tag;
  return (value)(b + 1);
}
static inline value caml_copy_double(double d) {
  value v = mock_caml_alloc(1, Double_tag);
  memcpy((void *)v, &d, sizeof d);
  return v;
}

/* exceptions */
extern jmp_buf mock_caml_handler;
extern char mock_caml_exn[512];
static inline void mock_caml_raise(const char *kind, const char *msg) {
  strncpy(mock_caml_exn, kind, sizeof mock_caml_exn - 1);
  strncat(mock_caml_exn, ": ", sizeof mock_caml_exn - strlen(mock_caml_exn) - 1);
  strncat(mock_caml_exn, msg ? msg : "", sizeof mock_caml_exn - strlen(mock_caml_exn) - 1);
  longjmp(mock_caml_handler, 1);
}
#endif
