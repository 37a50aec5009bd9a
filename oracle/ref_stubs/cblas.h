/* Stub standing in for ATLAS <cblas.h>; the kernel translation units include
 * it through NNTypes.h (VALIDATION) but call nothing from it. */
