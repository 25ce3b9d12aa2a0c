/* blow5_dump.c — fixture tooling (ours): dump every record of a BLOW5 file as
 *   [u32 id_len][id bytes][f64 digitisation][f64 offset][f64 range][f64 sampling_rate][u64 n][i16 raw * n]
 * Linked against the reference's vendored slow5lib, built from a scratch copy (see make_golden.py). */
#include <stdio.h>
#include <stdint.h>
#include <string.h>
#include <slow5/slow5.h>

int main(int argc, char** argv) {
    if (argc < 3) return 2;
    slow5_file_t* sf = slow5_open(argv[1], "r");
    if (!sf) return 1;
    FILE* out = fopen(argv[2], "wb");
    slow5_rec_t* rec = NULL;
    while (slow5_get_next(&rec, sf) >= 0) {
        uint32_t l = (uint32_t)strlen(rec->read_id);
        uint64_t n = rec->len_raw_signal;
        fwrite(&l, 4, 1, out);
        fwrite(rec->read_id, 1, l, out);
        fwrite(&rec->digitisation, 8, 1, out);
        fwrite(&rec->offset, 8, 1, out);
        fwrite(&rec->range, 8, 1, out);
        fwrite(&rec->sampling_rate, 8, 1, out);
        fwrite(&n, 8, 1, out);
        fwrite(rec->raw_signal, 2, n, out);
    }
    slow5_rec_free(rec);
    slow5_close(sf);
    fclose(out);
    return 0;
}
