// Host utilities with the call surface of the reference's src/data_utils.h (ReadFile :55, WriteFile :101,
// PrintData :147, printDataType :18-36, INFO/WARN/ERROR_LOG :38-40) so that a host written against the
// reference compiles against this header unchanged.  CHECK_ACL (:41-47) becomes CHECK_CUDA with the same
// report-and-continue behaviour; CHECK_ACL itself is kept as an alias.  The file functions are thin
// wrappers over the C ABI (ptb200_read_file / ptb200_write_file), which carries the reference's semantics.
#ifndef PTB200_DATA_UTILS_H
#define PTB200_DATA_UTILS_H
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <iomanip>
#include <iostream>
#include <string>

#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "../../include/ptb200.h"

// numeric tags are part of the interface (they match the reference's enum values)
typedef enum {
    DT_UNDEFINED = -1, FLOAT = 0, HALF = 1, INT8_T = 2, INT32_T = 3, UINT8_T = 4, INT16_T = 6, UINT16_T = 7, UINT32_T = 8,
    INT64_T = 9, UINT64_T = 10, DOUBLE = 11, BOOL = 12, STRING = 13, COMPLEX64 = 16, COMPLEX128 = 17, BF16 = 27
} printDataType;

// Same names, same output as the reference's log macros (one line on stdout: tag, two blanks, message).
namespace ptb200_detail {
__attribute__((format(printf, 2, 3))) inline void LogLine(const char *tag, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    fputs(tag, stdout);
    vfprintf(stdout, fmt, ap);
    fputc('\n', stdout);
    va_end(ap);
}
}  // namespace ptb200_detail
#define INFO_LOG(...) ptb200_detail::LogLine("[INFO]  ", __VA_ARGS__)
#define WARN_LOG(...) ptb200_detail::LogLine("[WARN]  ", __VA_ARGS__)
#define ERROR_LOG(...) ptb200_detail::LogLine("[ERROR]  ", __VA_ARGS__)

#define CHECK_CUDA(x)                                                                                               \
    do {                                                                                                            \
        cudaError_t ret__ = (x);                                                                                    \
        if (ret__ != cudaSuccess)                                                                                   \
            std::cerr << __FILE__ << ":" << __LINE__ << " cudaError:" << ret__ << " " << cudaGetErrorString(ret__) \
                      << std::endl;                                                                                 \
    } while (0)
#define CHECK_ACL(x) CHECK_CUDA(x)

#define CHECK_PTB200(x)                                                                                                       \
    do {                                                                                                                      \
        int ret__ = (x);                                                                                                      \
        if (ret__ != PTB200_OK)                                                                                               \
            std::cerr << __FILE__ << ":" << __LINE__ << " ptb200 error:" << ret__ << " " << ptb200_last_error() << std::endl; \
    } while (0)

inline bool ReadFile(const std::string &filePath, size_t &fileSize, void *buffer, size_t bufferSize) {
    size_t got = 0;
    if (ptb200_read_file(filePath.c_str(), &got, buffer, bufferSize) != PTB200_OK) {
        ERROR_LOG("%s", ptb200_last_error());
        return false;
    }
    fileSize = got;
    return true;
}

inline bool WriteFile(const std::string &filePath, const void *buffer, size_t size) {
    if (ptb200_write_file(filePath.c_str(), buffer, size) != PTB200_OK) {
        ERROR_LOG("%s", ptb200_last_error());
        return false;
    }
    return true;
}

namespace ptb200_detail {
template <typename T, typename AsT = T> void PrintRows(const void *data, size_t count, size_t perRow) {
    const T *v = static_cast<const T *>(data);
    for (size_t i = 0; i < count; ++i) {
        std::cout << std::setw(10) << static_cast<AsT>(v[i]);
        if ((i + 1) % perRow == 0)
            std::cout << std::endl;
    }
}
}  // namespace ptb200_detail

inline void PrintData(const void *data, size_t count, printDataType dataType, size_t elementsPerRow = 16) {
    if (data == nullptr) {
        ERROR_LOG("Print data failed. data is nullptr");
        return;
    }
    if (elementsPerRow == 0)
        elementsPerRow = 16;
    using namespace ptb200_detail;
    switch (dataType) {
    case BOOL: PrintRows<bool>(data, count, elementsPerRow); break;
    case INT8_T: PrintRows<int8_t>(data, count, elementsPerRow); break;
    case UINT8_T: PrintRows<uint8_t>(data, count, elementsPerRow); break;
    case INT16_T: PrintRows<int16_t>(data, count, elementsPerRow); break;
    case UINT16_T: PrintRows<uint16_t>(data, count, elementsPerRow); break;
    case INT32_T: PrintRows<int32_t>(data, count, elementsPerRow); break;
    case UINT32_T: PrintRows<uint32_t>(data, count, elementsPerRow); break;
    case INT64_T: PrintRows<int64_t>(data, count, elementsPerRow); break;
    case UINT64_T: PrintRows<uint64_t>(data, count, elementsPerRow); break;
    case FLOAT: PrintRows<float>(data, count, elementsPerRow); break;
    case DOUBLE: PrintRows<double>(data, count, elementsPerRow); break;
    case HALF: {
        const __half *v = static_cast<const __half *>(data);
        for (size_t i = 0; i < count; ++i) {
            std::cout << std::setw(10) << std::setprecision(6) << __half2float(v[i]);
            if ((i + 1) % elementsPerRow == 0)
                std::cout << std::endl;
        }
        break;
    }
    default: ERROR_LOG("Unsupported type: %d", dataType);
    }
    std::cout << std::endl;
}
#endif  // PTB200_DATA_UTILS_H
