#ifndef MOCK_CAML_THREADS_H
#define MOCK_CAML_THREADS_H
extern int mock_caml_runtime_released;  /* the driver checks that every release is paired with an acquire */
static inline void caml_release_runtime_system(void) { ++mock_caml_runtime_released; }
static inline void caml_acquire_runtime_system(void) { --mock_caml_runtime_released; }
#endif
