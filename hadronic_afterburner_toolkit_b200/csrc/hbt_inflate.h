// A gzip (RFC 1952 / deflate RFC 1951) decoder for the file reader (hbt_reader.cpp), written for the files this
// path reads: `%.17g`-style text, i.e. streams of literals with ~3.4 bits of entropy per character and few matches,
// where zlib's inflate (one table lookup and one bit-buffer refill per symbol through a state machine) delivers
// 140-350 MB/s of text and was the floor of the file-to-file path.  Same results as zlib for every valid
// stream (tests/test_inflate_cpu.py compares the two on text, binary, incompressible and highly compressible inputs,
// every block type, concatenated members, truncated and corrupted streams); errors where zlib reports errors.
//
// Host code only.  Not thread safe per stream; one stream per reader.
#ifndef HBT_INFLATE_H_
#define HBT_INFLATE_H_

#include <cstddef>
#include <cstdint>

struct HbtGz;  // opaque

extern "C" {
// nullptr when the file cannot be opened.  A file that does not start with the gzip magic is read as plain bytes
// (what gzread does for such files).
HbtGz *hbt_gz_open(const char *path);
// up to n decompressed bytes into buf; returns the number produced, 0 at the end of the file, -1 on error
long hbt_gz_read(HbtGz *gz, char *buf, size_t n);
const char *hbt_gz_error(const HbtGz *gz);
void hbt_gz_close(HbtGz *gz);
}

#endif  // HBT_INFLATE_H_
