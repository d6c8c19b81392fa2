// Test helper for csrc/cli/png_io.h:  png_roundtrip read|read4 <in.png> <out.raw>   (BGR8 / BGRA8 dump: int32 w, int32 h, bytes)
//                                     png_roundtrip write <in.raw> <out.png>  (raw = int32 w, h, channels, bytes)
#include <cstdio>
#include "../surround360_b200/csrc/cli/png_io.h"
int main(int argc, char** argv) {
  if (argc != 4) return 2;
  try {
    if (std::string(argv[1]) == "read" || std::string(argv[1]) == "read4") {
      pngio::Image im = pngio::readColor(argv[2], std::string(argv[1]) == "read4");
      FILE* f = fopen(argv[3], "wb");
      int32_t hdr[2] = {im.w, im.h};
      fwrite(hdr, 4, 2, f); fwrite(im.data.data(), 1, im.data.size(), f); fclose(f);
    } else {
      std::vector<uint8_t> b = pngio::readFile(argv[2]);
      int32_t hdr[3]; memcpy(hdr, b.data(), 12);
      pngio::write(argv[3], b.data() + 12, hdr[0], hdr[1], hdr[2]);
    }
  } catch (const std::exception& e) { fprintf(stderr, "%s\n", e.what()); return 1; }
  return 0;
}
