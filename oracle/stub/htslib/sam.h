/* Opaque stand-ins (see hts.h in this directory). */
#ifndef ORACLE_STUB_SAM_H
#define ORACLE_STUB_SAM_H
#include "hts.h"
typedef struct bam1_t bam1_t;
typedef struct bam_hdr_t bam_hdr_t;
#endif
