// predict -- top-K recommendations from a trained network (U/Predict.cpp:110-276, U/NNRecsGenerator.cpp:75-229), same
// arguments.  The exclusion filter (-f: a sample file listing what each sample already has, U/Filters.cpp:49-67) is
// applied on the device inside the top-K kernel instead of the reference's download / multiply-by-zero / upload.
#include <chrono>
#include <cstdio>
#include <iostream>

#include "../engine/NNNetwork.h"
#include "TextData.h"

using namespace textdata;
using std::cout;
using std::endl;

static void usage()
{
    cout << "Predict: Generates predictions from a trained neural network given a signals/input dataset." << endl;
    cout << "Usage: predict -d <dataset_name> -n <network_file> -r <input_text_file> -i <input_feature_index> -o <output_feature_index> -f <filters_json> [-b <batch_size>] [-k <num_recs>] [-l layer] [-s input_signals_index] [-p score_precision]" << endl;
    cout << "    -b batch_size: (default = 1024) the number records/input rows to process in a batch." << endl;
    cout << "    -d dataset_name: (required) name for the dataset within the netcdf file." << endl;
    cout << "    -f samples filterFileName ." << endl;
    cout << "    -i input_feature_index: (required) path to the feature index file, used to tranform input signals to correct input feature vector." << endl;
    cout << "    -k num_recs: (default = 100) The number of predictions (sorted by score to generate). Ignored if -l flag is used." << endl;
    cout << "    -l layer: (default = Output) the network layer to use for predictions. If specified, the raw scores for each node in the layer is output in order." << endl;
    cout << "    -n network_file: (required) the trained neural network in NetCDF file." << endl;
    cout << "    -o output_feature_index: (required) path to the feature index file, used to tranform the network output feature vector to appropriate features." << endl;
    cout << "    -p score_precision: (default = 4.3f) precision of the scores in output" << endl;
    cout << "    -r input_text_file: (required) path to the file with input signal to use to generate predictions (i.e. recommendations)." << endl;
    cout << "    -s filename (required) . to put the output recs to." << endl;
    cout << endl;
}

int main(int argc, char** argv)
{
    if (isArgSet(argc, argv, "-h")) { usage(); return 1; }
    const string dataSetName = getRequiredArgValue(argc, argv, "-d", "dataset_name is not specified.", &usage);
    const string filtersFileName = getRequiredArgValue(argc, argv, "-f", "filters_json is not specified.", &usage);
    if (!fileExists(filtersFileName)) { cout << "Error: Cannot read filter file: " << filtersFileName << endl; return 1; }
    const string inputIndexFileName = getRequiredArgValue(argc, argv, "-i", "input features index file is not specified.", &usage);
    if (!fileExists(inputIndexFileName)) { cout << "Error: Cannot read input feature index file: " << inputIndexFileName << endl; return 1; }
    const string networkFileName = getRequiredArgValue(argc, argv, "-n", "network file is not specified.", &usage);
    if (!fileExists(networkFileName)) { cout << "Error: Cannot read network file: " << networkFileName << endl; return 1; }
    const string outputIndexFileName = getRequiredArgValue(argc, argv, "-o", "output features index file is not specified.", &usage);
    if (!fileExists(outputIndexFileName)) { cout << "Error: Cannot read output feature index file: " << outputIndexFileName << endl; return 1; }
    const string recsFileName = getRequiredArgValue(argc, argv, "-r", "input_text_file is not specified.", &usage);
    if (!fileExists(recsFileName)) { cout << "Error: Cannot read input_text_file: " << recsFileName << endl; return 1; }
    const string recsOutputFileName = getRequiredArgValue(argc, argv, "-s", "filename to put the output recs to.", &usage);
    const unsigned int batchSize = (unsigned int)std::stoi(getOptionalArgValue(argc, argv, "-b", "1024"));
    const unsigned int topK = (unsigned int)std::stoi(getOptionalArgValue(argc, argv, "-k", "100"));
    if (topK >= 128) { cout << "Error :Optimized topk Only works for top 128 . " << topK << " is greater" << endl; return 1; }   // U/Predict.cpp:150-153
    const string scoreFormat = getOptionalArgValue(argc, argv, "-p", "4.3f");
    const string layer = getOptionalArgValue(argc, argv, "-l", "Output");

    try {
        getGpu().Startup(argc, argv);
        getGpu().SetRandomSeed(12134);
        if (getGpu()._numprocs > 1) throw DsbEngineError("predict: run with one process (the device-side top-K of this tool is single-GPU)");
        const auto preStart = std::chrono::steady_clock::now();
        Index mInput, mSignals, mOutput;
        cout << "Loading input feature index from: " << inputIndexFileName << endl;
        if (!loadIndexFromFile(mInput, inputIndexFileName, cout)) return 1;
        Csr signals;
        if (!parseSamples(recsFileName, false, mInput, mSignals, signals, cout)) return 1;       // convertTextToNetCDF, U/Predict.cpp:79-101
        if (signals.index.empty()) { cout << "Error: no known features in " << recsFileName << endl; return 1; }
        cout << "Loading output feature index from: " << outputIndexFileName << endl;
        if (!loadIndexFromFile(mOutput, outputIndexFileName, cout)) return 1;
        Index mFilterSamples = mSignals;
        Csr filter;
        if (!parseSamples(filtersFileName, false, mOutput, mFilterSamples, filter, cout)) return 1;
        const uint32_t examples = (uint32_t)signals.start.size();
        filter.start.resize(examples, filter.index.size());                                        // samples absent from the filter file: nothing excluded
        filter.end.resize(examples, filter.index.size());
        const vector<string> vSignals = invert(mSignals), vOutput = invert(mOutput);

        NNNetwork* pNetwork = LoadNeuralNetworkNetCDF(networkFileName, batchSize);
        // the input dataset, named after the network's input layer data set (<dataset_name>_input, U/Predict.cpp:186-189)
        const string inputName = dataSetName + "_input";
        const uint32_t width = roundUpMaxIndex((uint32_t)mInput.size());
        NNDataSet<uint32_t>* pInput = new NNDataSet<uint32_t>(examples, examples, signals.index.size(), NNDataSetDimensions(width), false, false, inputName);
        pInput->LoadSparseData(signals.start.data(), signals.end.data(), NULL, signals.index.data());
        pInput->_attributes |= NNDataSetEnums::Boolean;
        NNDataSet<uint32_t>* pFilter = new NNDataSet<uint32_t>(examples, examples, std::max<size_t>(filter.index.size(), 1), NNDataSetDimensions(roundUpMaxIndex((uint32_t)mOutput.size())),
                                                               false, false, "filter");
        if (filter.index.empty()) filter.index.push_back(0);
        pFilter->LoadSparseData(filter.start.data(), filter.end.data(), NULL, filter.index.data());
        vector<NNDataSetBase*> vDataSetInput(1, pInput);
        pNetwork->LoadDataSets(vDataSetInput);
        cout << "Total time for loading network and data is: " << std::chrono::duration<double>(std::chrono::steady_clock::now() - preStart).count() << endl;

        FILE* fp = fopen(recsOutputFileName.c_str(), "w");
        if (!fp) throw DsbEngineError("predict: cannot create " + recsOutputFileName);
        GpuBuffer<NNFloat> key((size_t)batchSize * topK, true);
        GpuBuffer<uint32_t> value((size_t)batchSize * topK, true);
        const string fmt = "%s,%" + scoreFormat + ":";
        const auto recsStart = std::chrono::steady_clock::now();
        for (uint64_t pos = 0; pos < pNetwork->GetExamples(); pos += pNetwork->GetBatch()) {
            cout << "Predicting from position " << pos << endl;
            pNetwork->SetPosition((uint32_t)pos);
            pNetwork->PredictBatch();
            pNetwork->CalculateTopKFiltered(layer, topK, pFilter, &key, &value);
            key.Download();
            value.Download();
            const uint32_t batch = (uint32_t)std::min<uint64_t>(pNetwork->GetBatch(), pNetwork->GetExamples() - pos);
            for (uint32_t j = 0; j < batch; j++) {
                fprintf(fp, "%s\t", vSignals[pos + j].c_str());
                for (uint32_t x = 0; x < topK; x++) {
                    const uint32_t f = value._pSysData[(size_t)j * topK + x];
                    if (f < vOutput.size()) fprintf(fp, fmt.c_str(), vOutput[f].c_str(), key._pSysData[(size_t)j * topK + x]);
                }
                fprintf(fp, "\n");
            }
        }
        fclose(fp);
        cout << "Total time for Generating recs for " << pNetwork->GetExamples() << " was "
             << std::chrono::duration<double>(std::chrono::steady_clock::now() - recsStart).count() << endl;
        delete pNetwork;
        delete pInput;
        delete pFilter;
        getGpu().Shutdown();
    } catch (const std::exception& e) {
        cout << "Error: " << e.what() << endl;
        return 1;
    }
    return 0;
}
