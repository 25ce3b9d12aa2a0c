/* blow5_recompress.c — fixture tooling (ours): rewrite a BLOW5 file with another record / signal compression using the
 * reference's vendored slow5lib (slow5_convert), so that the fixtures of the device-side decoders (svb-zd signal
 * compression, slow5lib/src/slow5_press.c:1063-1170) are produced by the reference's own encoder.
 *   blow5_recompress in.blow5 out.blow5 <record: none|zlib> <signal: none|svb-zd|ex-zd>
 * Linked against slow5lib built from a scratch copy (see make_blow5_fixtures.py). */
#include <stdio.h>
#include <string.h>
#include <slow5/slow5.h>

int main(int argc, char** argv) {
    if (argc < 5) return 2;
    slow5_file_t* sf = slow5_open(argv[1], "r");
    if (!sf) return 1;
    FILE* out = fopen(argv[2], "wb");
    if (!out) return 1;
    slow5_press_method_t m;
    m.record_method = strcmp(argv[3], "zlib") == 0 ? SLOW5_COMPRESS_ZLIB : SLOW5_COMPRESS_NONE;
    m.signal_method = strcmp(argv[4], "svb-zd") == 0 ? SLOW5_COMPRESS_SVB_ZD
                      : (strcmp(argv[4], "ex-zd") == 0 ? SLOW5_COMPRESS_EX_ZD : SLOW5_COMPRESS_NONE);
    int rc = slow5_convert(sf, out, SLOW5_FORMAT_BINARY, m);
    if (rc == 0) { /* slow5_convert writes header + records; the end-of-file marker is the caller's */
        const char eof[] = {'5', 'W', 'O', 'L', 'B'};
        fwrite(eof, 1, sizeof eof, out);
    }
    fclose(out);
    slow5_close(sf);
    return rc;
}
