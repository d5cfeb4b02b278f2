/* fur_reader.h -- reference-built .fur/.mfur file -> flattened device image (image.h). Host only. */
#ifndef FULGOR_B200_FUR_READER_H
#define FULGOR_B200_FUR_READER_H

#include <cstdint>
#include <vector>

#include "image.h"

namespace fgb {

/* 0 = hybrid (.fur), 1 = meta (.mfur), -2 = known but unsupported (.dfur/.mdfur), -1 = unknown;
   by suffix, like reference tools/util.cpp:5-19 */
int index_type_from_path(const char* path);

/* throw std::runtime_error on malformed / unsupported input */
std::vector<uint8_t> build_image(const uint8_t* file_bytes, uint64_t size, int type);
std::vector<uint8_t> build_image_from_file(const char* path);

}  // namespace fgb
#endif
