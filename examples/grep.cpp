// grep.cpp -- B200 backend for the reference's example CLI (examples/grep.rs:42-57):
//   ./grep <backend> <needle> <file>      backend: b200 | dynamicb200
// mmaps the file and runs one search_in over the host slice (ss_b200_search_in_host: chunked
// host->device streaming overlapped with the scan), printing the same line as the reference.
#include "sliceslice_b200.hpp"

#include <cstdio>
#include <cstring>
#include <fcntl.h>
#include <strings.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

using namespace sliceslice_b200;

// examples/grep.rs:12-40
static bool search_in_slice(const char *backend, Bytes needle, Bytes haystack)
{
    if (!strcasecmp(backend, "b200"))
        return B200Searcher::new_(needle).search_in(haystack);
    if (!strcasecmp(backend, "dynamicb200"))
        return DynamicB200Searcher::new_(needle).search_in(haystack);
    fprintf(stderr, "Invalid backend \"%s\"\n", backend);
    exit(101);
}

int main(int argc, char **argv)
{
    if (argc != 4) {
        fprintf(stderr, "./grep <backend> <needle> <file>\n");
        return 101;
    }
    const char *backend = argv[1], *needle = argv[2], *filename = argv[3];
    const int fd = open(filename, O_RDONLY);
    struct stat st;
    if (fd < 0 || fstat(fd, &st) != 0) {
        perror(filename);
        return 101;
    }
    const void *data = st.st_size ? mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0) : "";
    if (data == MAP_FAILED) {
        perror("mmap");
        return 101;
    }
    try {
        const bool found = search_in_slice(backend, Bytes(needle), Bytes(data, (size_t)st.st_size));
        printf("Searching for %s in \"%s\": %s\n", needle, filename, found ? "true" : "false");
    } catch (const std::exception &e) { // the reference panics (exit code 101)
        fprintf(stderr, "panicked: %s\n", e.what());
        return 101;
    }
    return 0;
}
