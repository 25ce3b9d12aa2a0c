/* Opaque stand-ins for the htslib types that the reference's f5c.h names but the ABEA path never
 * dereferences. Test infrastructure only (oracle build); htslib itself is absent from this image. */
#ifndef ORACLE_STUB_HTS_H
#define ORACLE_STUB_HTS_H
typedef struct htsFile htsFile;
typedef struct hts_idx_t hts_idx_t;
typedef struct hts_itr_t hts_itr_t;
typedef htsFile samFile;
#endif
