#ifndef MOCK_CAML_MEMORY_H
#define MOCK_CAML_MEMORY_H
#include "mlvalues.h"
/* no GC: roots need no registration */
#define CAMLparam0() (void)0
#define CAMLparam1(a) (void)(a)
#define CAMLparam2(a, b) (void)(a), (void)(b)
#define CAMLparam3(a, b, c) (void)(a), (void)(b), (void)(c)
#define CAMLparam4(a, b, c, d) (void)(a), (void)(b), (void)(c), (void)(d)
#define CAMLparam5(a, b, c, d, e) (void)(a), (void)(b), (void)(c), (void)(d), (void)(e)
#define CAMLxparam1(a) (void)(a)
#define CAMLxparam2(a, b) (void)(a), (void)(b)
#define CAMLxparam3(a, b, c) (void)(a), (void)(b), (void)(c)
#define CAMLlocal1(a) value a = Val_unit
#define CAMLlocal2(a, b) value a = Val_unit, b = Val_unit
#define CAMLreturn(x) return (x)
#define CAMLreturn0 return
static inline void *caml_stat_alloc(size_t n) { return malloc(n ? n : 1); }
static inline void caml_stat_free(void *p) { free(p); }
#endif
