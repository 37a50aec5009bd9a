// train -- trains a network from a JSON config and NetCDF datasets (U/Train.cpp:59-168), same arguments and defaults.
// Launch one process per GPU (torchrun-style RANK / WORLD_SIZE / LOCAL_RANK in the environment) for model-parallel runs.
#include <chrono>
#include <iostream>

#include "../engine/NNNetwork.h"
#include "TextData.h"

using namespace textdata;
using std::cout;
using std::endl;

static void usage()
{
    cout << "Train: Trains a neural networks given a config and dataset." << endl;
    cout << "Usage: train -d <dataset_name> -c <config_file> -n <network_file> -i <input_netcdf> -o <output_netcdf> [-b <batch_size>] [-e <num_epochs>]" << endl;
    cout << "    -c config_file: (required) the JSON config files with network training parameters." << endl;
    cout << "    -i input_netcdf: (required) path to the netcdf with dataset for the input of the network." << endl;
    cout << "    -o output_netcdf: (required) path to the netcdf with dataset for expected output of the network." << endl;
    cout << "    -n network_file: (required) the output trained neural network in NetCDF file." << endl;
    cout << "    -b batch_size: (default = 1024) the number records/input rows to process in a batch." << endl;
    cout << "    -e num_epochs: (default = 40) the number passes on the full dataset." << endl;
    cout << "    -m mode: (default = SGD) SGD | Momentum | AdaGrad | Nesterov | RMSProp | AdaDelta | Adam (B200 addition)." << endl;
    cout << "    -g gemm_mode: (default = 2) dense GEMMs: 0 cuBLAS fp32, 1 tcgen05 TF32, 2 tcgen05 3xTF32 (B200 addition)." << endl;
    cout << endl;
}

int main(int argc, char** argv)
{
    const float alpha = std::stof(getOptionalArgValue(argc, argv, "-alpha", "0.025f"));
    const float lambda = std::stof(getOptionalArgValue(argc, argv, "-lambda", "0.0001f"));
    const float lambda1 = std::stof(getOptionalArgValue(argc, argv, "-lambda1", "0.0f"));
    const float mu = std::stof(getOptionalArgValue(argc, argv, "-mu", "0.5f"));
    const float mu1 = std::stof(getOptionalArgValue(argc, argv, "-mu1", "0.0f"));
    if (isArgSet(argc, argv, "-h")) { usage(); return 1; }
    const string configFileName = getRequiredArgValue(argc, argv, "-c", "config file was not specified.", &usage);
    if (!fileExists(configFileName)) { cout << "Error: Cannot read config file: " << configFileName << endl; return 1; }
    cout << "Train will use configuration file: " << configFileName << endl;
    const string inputDataFile = getRequiredArgValue(argc, argv, "-i", "input data file is not specified.", &usage);
    if (!fileExists(inputDataFile)) { cout << "Error: Cannot read input feature index file: " << inputDataFile << endl; return 1; }
    cout << "Train will use input data file: " << inputDataFile << endl;
    const string outputDataFile = getRequiredArgValue(argc, argv, "-o", "output data  file is not specified.", &usage);
    if (!fileExists(outputDataFile)) { cout << "Error: Cannot read output feature index file: " << outputDataFile << endl; return 1; }
    cout << "Train will use output data file: " << outputDataFile << endl;
    const string networkFileName = getRequiredArgValue(argc, argv, "-n", "the output network file path is not specified.", &usage);
    if (fileExists(networkFileName)) { cout << "Error: Network file already exists: " << networkFileName << endl; return 1; }
    cout << "Train will produce networkFileName: " << networkFileName << endl;
    const unsigned int batchSize = (unsigned int)std::stoi(getOptionalArgValue(argc, argv, "-b", "1024"));
    cout << "Train will use batchSize: " << batchSize << endl;
    const unsigned int epoch = (unsigned int)std::stoi(getOptionalArgValue(argc, argv, "-e", "40"));
    cout << "Train will use number of epochs: " << epoch << endl;
    cout << "Train alpha " << alpha << ", lambda " << lambda << ", lambda1 " << lambda1 << ", mu " << mu << ", mu1 " << mu1 << ".Please check CDL.txt for meanings" << endl;
    const string modeName = getOptionalArgValue(argc, argv, "-m", "SGD");
    static const char* names[] = {"SGD", "Momentum", "AdaGrad", "Nesterov", "RMSProp", "AdaDelta", "Adam"};
    int mode = -1;
    for (int i = 0; i < 7; i++)
        if (modeName == names[i]) mode = i;
    if (mode < 0) { cout << "Error: unknown training mode " << modeName << endl; return 1; }
    const int gemmMode = std::stoi(getOptionalArgValue(argc, argv, "-g", "2"));

    try {
        getGpu().Startup(argc, argv);
        getGpu().SetRandomSeed(12134);                                        // FIXED_SEED, U/Utils.h:27
        dsb200_ctx_set_option(getGpu()._ctx, "gemm_mode", gemmMode);
        vector<NNDataSetBase*> vDataSetInput = LoadNetCDF(inputDataFile);
        vector<NNDataSetBase*> vDataSetOutput = LoadNetCDF(outputDataFile);
        vDataSetInput.insert(vDataSetInput.end(), vDataSetOutput.begin(), vDataSetOutput.end());
        NNNetwork* pNetwork = LoadNeuralNetworkJSON(configFileName, batchSize, vDataSetInput);
        pNetwork->LoadDataSets(vDataSetInput);
        pNetwork->SetCheckpoint(networkFileName, 10);
        pNetwork->SetPosition(0);
        pNetwork->PredictBatch();
        pNetwork->SaveNetCDF("initial_network.nc");
        pNetwork->SetTrainingMode((TrainingMode)mode);
        const auto start = std::chrono::steady_clock::now();
        for (unsigned int x = 0; x < epoch; ++x) {
            const float error = pNetwork->Train(1, alpha, lambda, lambda1, mu, mu1);
            if (getGpu()._id == 0) cout << "Epoch " << x + 1 << " Average_Error " << error << endl;
        }
        getGpu().Synchronize();
        const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - start).count();
        cout << "Total Training Time " << secs << endl;
        int totalGPUMemory, totalCPUMemory;
        getGpu().GetMemoryUsage(&totalGPUMemory, &totalCPUMemory);
        cout << "GPU Memory Usage: " << totalGPUMemory << " KB" << endl;
        cout << "CPU Memory Usage: " << totalCPUMemory << " KB" << endl;
        pNetwork->SaveNetCDF(networkFileName);
        delete pNetwork;
        for (auto p : vDataSetInput) delete p;
        getGpu().Shutdown();
    } catch (const std::exception& e) {
        cout << "Error: " << e.what() << endl;
        return 1;
    }
    return 0;
}
