// generateNetCDF -- text samples -> NetCDF dataset + index files (U/NetCDFGenerator.cpp:30-150), host only.
#include <chrono>
#include <iostream>

#include "../../include/dsstne_b200_engine.h"
#include "TextData.h"

using namespace std;
using namespace textdata;

static void usage()
{
    cout << "NetCDFGenerator: Converts a text dataset file into a more compressed NetCDF file." << endl;
    cout << "Usage: generateNetCDF -d <dataset_name> -i <input_text_file> -o <output_netcdf_file> -f <features_index> -s <samples_index> [-c] [-m]" << endl;
    cout << "    -d dataset_name: (required) name for the dataset within the netcdf file." << endl;
    cout << "    -i input_text_file: (required) path to the input text file with records in data format." << endl;
    cout << "    -o output_netcdf_file: (required) path to the output netcdf file that we generate." << endl;
    cout << "    -f features_index: (required) path to the features index file to read-from/write-to." << endl;
    cout << "    -s samples_index: (required) path to the samples index file to read-from/write-to." << endl;
    cout << "    -m : if set, we'll merge the feature index with new features found in the input_text_file. (Cannot be used with -c)." << endl;
    cout << "    -c : if set, we'll create a new feature index from scratch. (Cannot be used with -m)." << endl;
    cout << "    -t type: (default = indicator) the type of dataset to generate. Valid values are: ['indicator', 'analog']." << endl;
    cout << endl;
}

int main(int argc, char** argv)
{
    if (isArgSet(argc, argv, "-h")) { usage(); return 1; }
    const string inputFile = getRequiredArgValue(argc, argv, "-i", "input text file to convert.", &usage);
    const string outputFile = getRequiredArgValue(argc, argv, "-o", "output netcdf file to generate.", &usage);
    const string datasetName = getRequiredArgValue(argc, argv, "-d", "dataset name for the netcdf metadata.", &usage);
    const string featureIndexFile = getRequiredArgValue(argc, argv, "-f", "feature index file.", &usage);
    const string sampleIndexFile = getRequiredArgValue(argc, argv, "-s", "samples index file.", &usage);
    const bool create = isArgSet(argc, argv, "-c"), merge = isArgSet(argc, argv, "-m");
    if (create && merge) { cout << "Error: Cannot create (-c) and update existing (-u) feature index. Please select only one." << endl; return 1; }
    const string dataType = getOptionalArgValue(argc, argv, "-t", "indicator");
    if (dataType != "indicator" && dataType != "analog") {
        cout << "Error: Unknown dataset type [" << dataType << "]. Please select one of {indicator,analog}" << endl;
        return 1;
    }
    cout << "Generating dataset of type: " << dataType << endl;
    const auto start = chrono::steady_clock::now();
    Index features, samples;
    if (fileExists(sampleIndexFile)) {
        cout << "Loading sample index from: " << sampleIndexFile << endl;
        if (!loadIndexFromFile(samples, sampleIndexFile, cout)) return 1;
    } else cout << "Will create a new samples index file: " << sampleIndexFile << endl;
    if (create) cout << "Will create a new features index file: " << featureIndexFile << endl;
    else if (!fileExists(featureIndexFile)) { cout << "Error: Cannnot find a valid feature index file: " << featureIndexFile << endl; return 1; }
    else {
        cout << "Loading feature index from: " << featureIndexFile << endl;
        if (!loadIndexFromFile(features, featureIndexFile, cout)) return 1;
    }
    const size_t samplesBefore = samples.size(), featuresBefore = features.size();
    Csr csr;
    if (!parseSamples(inputFile, create || merge, features, samples, csr, cout)) return 1;
    if (features.size() != featuresBefore) { exportIndex(features, featureIndexFile); cout << "Exported " << featureIndexFile << " with " << features.size() << " entries." << endl; }
    if (samples.size() != samplesBefore) { exportIndex(samples, sampleIndexFile); cout << "Exported " << sampleIndexFile << " with " << samples.size() << " entries." << endl; }
    if (csr.index.empty()) { cout << "Error: no data points found in " << inputFile << endl; return 1; }
    const unsigned int width = roundUpMaxIndex((unsigned int)features.size());
    cout << "Raw max index is: " << features.size() << endl << "Rounded up max index to: " << width << endl;
    const bool analog = dataType == "analog";
    const int rc = dsb200_netcdf_write_sparse(outputFile.c_str(), 5, datasetName.c_str(), 0, analog ? 4 /*Float*/ : 0 /*UInt*/, width, (uint32_t)csr.start.size(),
                                              (uint32_t)csr.start.size(), csr.start.data(), csr.end.data(), csr.index.data(), analog ? csr.data.data() : NULL,
                                              NULL, NULL);
    if (rc) { cout << "Error writing to NetCDF file: " << dsb200_engine_last_error() << endl; return 1; }
    cout << "Created NetCDF file " << outputFile << " for dataset " << datasetName << endl;
    cout << "Total time for generating NetCDF: " << chrono::duration<double>(chrono::steady_clock::now() - start).count() << " secs. " << endl;
    return 0;
}
