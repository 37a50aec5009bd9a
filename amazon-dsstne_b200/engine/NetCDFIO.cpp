// NetCDFIO.cpp -- dataset / network NetCDF files ("next" row 1 of SURVEY 8f).  PLACEHOLDER:
// the classic-format reader/writer lands after the hot path is measured; until then these entry
// points fail loudly instead of pretending.
#include "NNNetwork.h"

vector<NNDataSetBase*> LoadNetCDF(const string& fname)
{
    throw DsbEngineError("LoadNetCDF(" + fname + "): NetCDF dataset files are not built yet (SURVEY 8f row 1)");
}

bool SaveNetCDF(const string& fname, vector<NNDataSetBase*> vDataSet)
{
    (void)vDataSet;
    throw DsbEngineError("SaveNetCDF(" + fname + "): NetCDF dataset files are not built yet (SURVEY 8f row 1)");
}

bool NNNetwork::SaveNetCDF(const string& fname)
{
    throw DsbEngineError("NNNetwork::SaveNetCDF(" + fname + "): NetCDF network files are not built yet (SURVEY 8f row 1)");
}

NNNetwork* LoadNeuralNetworkNetCDF(const string& fname, const uint32_t batch)
{
    (void)batch;
    throw DsbEngineError("LoadNeuralNetworkNetCDF(" + fname + "): NetCDF network files are not built yet (SURVEY 8f row 1)");
}
