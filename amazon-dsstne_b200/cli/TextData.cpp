// TextData.cpp -- see TextData.h.
#include "TextData.h"

#include <cstdlib>
#include <fstream>
#include <sstream>

namespace textdata {

static std::vector<std::string> split(const std::string& s, char delim)
{
    std::vector<std::string> out;
    std::string item;
    std::stringstream ss(s);
    while (std::getline(ss, item, delim)) out.push_back(item);
    return out;
}

bool loadIndexFromFile(Index& index, const std::string& fname, std::ostream& log)
{
    std::ifstream in(fname);
    if (!in) { log << "Error: Failed to open index file " << fname << std::endl; return false; }
    std::string line;
    unsigned int lines = 0;
    const size_t before = index.size();
    while (std::getline(in, line)) {
        lines++;
        const std::vector<std::string> v = split(line, '\t');
        if (v.size() == 2 && !v[0].empty()) index[v[0]] = (unsigned int)atoi(v[1].c_str());
        else { log << "Error: line " << lines << " contains invalid data" << std::endl; return false; }
    }
    log << "Number of lines processed: " << lines << std::endl;
    log << "Number of entries added to index: " << index.size() - before << std::endl;
    if (lines != index.size() - before) {
        log << "Error: Number of entries added to index not equal to number of lines processed" << std::endl;
        return false;
    }
    return true;
}

void exportIndex(const Index& index, const std::string& fname)
{
    std::ofstream out(fname);
    for (const auto& kv : index) out << kv.first << "\t" << kv.second << std::endl;
}

std::vector<std::string> invert(const Index& index)
{
    std::vector<std::string> v(index.size());
    for (const auto& kv : index)
        if (kv.second < v.size()) v[kv.second] = kv.first;
    return v;
}

bool parseSamples(const std::string& fname, bool growFeatures, Index& features, Index& samples, Csr& out, std::ostream& log)
{
    std::ifstream in(fname);
    if (!in) { log << "Error: Failed to open sample file " << fname << std::endl; return false; }
    // rows keyed by sample index; a sample may appear on several lines (its features accumulate)
    std::vector<std::vector<std::pair<uint32_t, float>>> rows(samples.size());
    std::string line;
    size_t lineNo = 0;
    while (std::getline(in, line)) {
        lineNo++;
        if (line.empty()) continue;
        const size_t tab = line.find('\t');
        if (tab == std::string::npos) { log << "Warning: line " << lineNo << " has no tab, skipped" << std::endl; continue; }
        const std::string sample = line.substr(0, tab);
        if (sample.empty()) { log << "Warning: line " << lineNo << " has an empty sample name, skipped" << std::endl; continue; }
        auto it = samples.find(sample);
        unsigned int row;
        if (it == samples.end()) { row = (unsigned int)samples.size(); samples[sample] = row; rows.resize(row + 1); }
        else { row = it->second; if (row >= rows.size()) rows.resize(row + 1); }
        for (const std::string& item : split(line.substr(tab + 1), ':')) {
            if (item.empty()) continue;
            const size_t comma = item.find(',');
            const std::string feat = item.substr(0, comma);
            if (feat.empty()) continue;
            float value = 1.0f;
            if (comma != std::string::npos && comma + 1 < item.size()) value = (float)atof(item.c_str() + comma + 1);
            auto fit = features.find(feat);
            unsigned int col;
            if (fit == features.end()) {
                if (!growFeatures) continue;
                col = (unsigned int)features.size();
                features[feat] = col;
            } else col = fit->second;
            rows[row].push_back(std::make_pair((uint32_t)col, value));
        }
    }
    out.start.assign(rows.size(), 0);
    out.end.assign(rows.size(), 0);
    out.index.clear();
    out.data.clear();
    for (size_t r = 0; r < rows.size(); r++) {
        out.start[r] = out.index.size();
        for (const auto& e : rows[r]) { out.index.push_back(e.first); out.data.push_back(e.second); }
        out.end[r] = out.index.size();
    }
    log << "Number of samples: " << rows.size() << ", data points: " << out.index.size() << ", features: " << features.size() << std::endl;
    return true;
}

bool isArgSet(int argc, char** argv, const std::string& flag)
{
    for (int i = 1; i < argc; i++)
        if (flag == argv[i]) return true;
    return false;
}

std::string getOptionalArgValue(int argc, char** argv, const std::string& flag, const std::string& dflt)
{
    for (int i = 1; i + 1 < argc; i++)
        if (flag == argv[i]) return argv[i + 1];
    return dflt;
}

std::string getRequiredArgValue(int argc, char** argv, const std::string& flag, const std::string& message, void (*usage)())
{
    for (int i = 1; i + 1 < argc; i++)
        if (flag == argv[i]) return argv[i + 1];
    std::cout << "Error: Missing required argument: " << flag << ": " << message << std::endl;
    usage();
    exit(1);
}

bool fileExists(const std::string& fname)
{
    std::ifstream f(fname);
    return f.good();
}

}  // namespace textdata
