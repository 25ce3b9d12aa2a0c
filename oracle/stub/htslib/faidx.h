/* Opaque stand-ins (see hts.h in this directory). */
#ifndef ORACLE_STUB_FAIDX_H
#define ORACLE_STUB_FAIDX_H
typedef struct faidx_t faidx_t;
#endif
