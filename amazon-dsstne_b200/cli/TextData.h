// TextData.h -- the text side of DSSTNE's tools: index files (`label<TAB>index` per line) and sample files
// (`sample<TAB>feature[,value]:feature[,value]:...` per line), as U/NetCDFhelper.cpp:36-323 reads and writes them.
#pragma once

#include <cstdint>
#include <iostream>
#include <string>
#include <unordered_map>
#include <vector>

namespace textdata {

typedef std::unordered_map<std::string, unsigned int> Index;

// label<TAB>index lines -> map; false (with a message on `log`) on malformed or duplicate lines (loadIndex, U/NetCDFhelper.cpp:36-67)
bool loadIndexFromFile(Index& index, const std::string& fname, std::ostream& log);
void exportIndex(const Index& index, const std::string& fname);
// index -> label table (extractNNMapsToVectors, U/Predict.cpp:46-55)
std::vector<std::string> invert(const Index& index);

struct Csr {
    std::vector<uint64_t> start, end;
    std::vector<uint32_t> index;
    std::vector<float> data;             // value after the comma (1 when absent); what "analog" datasets keep
};

// Parses a sample file.  Unknown samples are appended to `samples` in order of first appearance; unknown features are
// appended to `features` when `growFeatures`, otherwise skipped (as the reference does at prediction time).  Rows are
// emitted in SAMPLE INDEX order, one row per known sample (empty when the file has no line for it).
bool parseSamples(const std::string& fname, bool growFeatures, Index& features, Index& samples, Csr& out, std::ostream& log);

// maxFeatureIndex rounded up to a multiple of 128 (roundUpMaxIndex, U/NetCDFhelper.cpp:325-330)
inline unsigned int roundUpMaxIndex(unsigned int n) { return ((n + 127) >> 7) << 7; }

// command line helpers with the reference's behaviour (U/Utils.cpp:48-113)
bool isArgSet(int argc, char** argv, const std::string& flag);
std::string getOptionalArgValue(int argc, char** argv, const std::string& flag, const std::string& dflt);
std::string getRequiredArgValue(int argc, char** argv, const std::string& flag, const std::string& message, void (*usage)());
bool fileExists(const std::string& fname);

}  // namespace textdata
