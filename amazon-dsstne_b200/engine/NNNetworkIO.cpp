// NNNetworkIO.cpp -- network description loaders.
//
// LoadNeuralNetworkJSON parses the reference's JSON layer-description language ("LDL",
// docs/getting_started/LDL.txt; parser E/NNNetwork.cpp:2792-3759) for the fully-connected subset:
// case-insensitive keys, an unknown key is fatal, same defaults (E/NNNetwork.cpp:27-58,
// E/NNLayer.cpp:2963-2998).  jsoncpp is not available offline, so a ~120-line recursive-descent
// JSON reader is included.  Keys that select layer types outside the hot path (Convolutional,
// Pooling, BatchNormalization, SharedWeights, Skip, LRN, Maxout) are recognised and rejected with
// an explicit message rather than silently ignored.
#include <algorithm>
#include <cctype>
#include <fstream>
#include <map>
#include <set>
#include <sstream>

#include "NNNetwork.h"

using namespace std;

namespace {

struct JValue {
    enum Type { Null, Bool, Number, String, Array, Object } type = Null;
    bool b = false;
    double num = 0.0;
    string str;
    vector<JValue> arr;
    vector<pair<string, JValue>> obj;          // insertion order kept (the reference iterates members in order)
    bool isArray() const { return type == Array; }
    bool isString() const { return type == String; }
    bool isObject() const { return type == Object; }
    double asNumber() const
    {
        if (type == Number) return num;
        if (type == Bool) return b ? 1.0 : 0.0;
        if (type == String) return atof(str.c_str());
        throw DsbEngineError("LoadNeuralNetworkJSON: expected a number");
    }
    bool asBool() const { return type == Bool ? b : (type == Number ? num != 0.0 : (type == String ? (str == "true") : false)); }
    string asString() const
    {
        if (type == String) return str;
        if (type == Number) { ostringstream o; o << num; return o.str(); }
        if (type == Bool) return b ? "true" : "false";
        return "";
    }
};

struct JParser {
    const string& s;
    size_t p = 0;
    explicit JParser(const string& text) : s(text) {}
    [[noreturn]] void fail(const string& what) { throw DsbEngineError("LoadNeuralNetworkJSON: JSON parse error at offset " + to_string(p) + ": " + what); }
    void ws() { while (p < s.size() && isspace((unsigned char)s[p])) p++; }
    JValue parse()
    {
        ws();
        if (p >= s.size()) fail("unexpected end");
        const char c = s[p];
        JValue v;
        if (c == '{') {
            v.type = JValue::Object; p++; ws();
            if (p < s.size() && s[p] == '}') { p++; return v; }
            for (;;) {
                ws();
                if (p >= s.size() || s[p] != '"') fail("expected a member name");
                string key = parseString();
                ws();
                if (p >= s.size() || s[p] != ':') fail("expected ':'");
                p++;
                v.obj.push_back(make_pair(key, parse()));
                ws();
                if (p < s.size() && s[p] == ',') { p++; continue; }
                if (p < s.size() && s[p] == '}') { p++; break; }
                fail("expected ',' or '}'");
            }
        } else if (c == '[') {
            v.type = JValue::Array; p++; ws();
            if (p < s.size() && s[p] == ']') { p++; return v; }
            for (;;) {
                v.arr.push_back(parse());
                ws();
                if (p < s.size() && s[p] == ',') { p++; continue; }
                if (p < s.size() && s[p] == ']') { p++; break; }
                fail("expected ',' or ']'");
            }
        } else if (c == '"') {
            v.type = JValue::String; v.str = parseString();
        } else if (!s.compare(p, 4, "true")) { v.type = JValue::Bool; v.b = true; p += 4; }
        else if (!s.compare(p, 5, "false")) { v.type = JValue::Bool; v.b = false; p += 5; }
        else if (!s.compare(p, 4, "null")) { v.type = JValue::Null; p += 4; }
        else {
            char* end = NULL;
            v.num = strtod(s.c_str() + p, &end);
            if (end == s.c_str() + p) fail("unexpected character");
            v.type = JValue::Number;
            p = end - s.c_str();
        }
        return v;
    }
    string parseString()
    {
        string out; p++;
        while (p < s.size() && s[p] != '"') {
            if (s[p] == '\\' && p + 1 < s.size()) {
                p++;
                switch (s[p]) { case 'n': out += '\n'; break; case 't': out += '\t'; break; case 'r': out += '\r'; break; default: out += s[p]; }
            } else out += s[p];
            p++;
        }
        if (p >= s.size()) fail("unterminated string");
        p++;
        return out;
    }
};

string lower(string s) { transform(s.begin(), s.end(), s.begin(), ::tolower); return s; }

[[noreturn]] void bad(const string& what) { throw DsbEngineError("LoadNeuralNetworkJSON: " + what); }

Activation parseActivation(const string& v)
{
    const string s = lower(v);                                  // E/NNNetwork.cpp:3418-3440
    if (s == "sigmoid") return Sigmoid;
    if (s == "tanh") return Tanh;
    if (s == "linear") return Linear;
    if (s == "relu" || s == "rectifiedlinear") return RectifiedLinear;
    if (s == "lrelu" || s == "leakyrectifiedlinear") return LeakyRectifiedLinear;
    if (s == "elu" || s == "exponentiallinear") return ExponentialLinear;
    if (s == "selu" || s == "scaledexponentiallinear") return ScaledExponentialLinear;
    if (s == "softmax") return SoftMax;
    bad("Invalid or unsupported layer activation: " + v + " (supported on this path: Sigmoid, Tanh, Linear, RELU, LRELU, ELU, SELU, SoftMax)");
}

WeightInitialization parseScheme(const string& v)
{
    const string s = lower(v);                                  // E/NNNetwork.cpp:3467-3488
    if (s == "xavier") return Xavier;
    if (s == "caffexavier") return CaffeXavier;
    if (s == "gaussian") return Gaussian;
    if (s == "uniform") return Uniform;
    if (s == "unitball") return UnitBall;
    if (s == "constant") return Constant;
    if (s == "selu") return SELU;
    bad("Invalid weight initialization scheme: " + v);
}

}  // namespace

// The JSON text -> NNNetworkDescriptor stage needs no GPU: the only thing it takes from the data sets is their dimensions
// (auto-sized input / output layers, E/NNNetwork.cpp:3600-3622).  DescribeNeuralNetworkJSON runs it alone, so that the
// reference's own sample configurations can be checked on a machine without a device.
NNNetworkDescriptor ParseNeuralNetworkJSON(const string& json, const vector<NNDataSetShape>& vDataSet)
{
    JParser parser(json);
    JValue root = parser.parse();
    if (!root.isObject()) bad("top level must be an object");
    NNNetworkDescriptor nd;
    set<string> sLayer;
    float version = NN_VERSION;
    for (auto& m : root.obj) {
        const string name = lower(m.first);
        const JValue& value = m.second;
        if (name == "version") {
            version = (float)value.asNumber();
            if (version < 0.6999f) bad("version must be >= 0.7");
        } else if (name == "name") nd._name = value.asString();
        else if (name == "kind") {
            const string s = lower(value.asString());
            if (s == "feedforward") nd._kind = NNNetwork::Kind::FeedForward;
            else if (s == "autoencoder") nd._kind = NNNetwork::Kind::AutoEncoder;
            else bad("Invalid network kind: " + value.asString());
        } else if (name == "weightsdata") bad("WeightsData (external weight file) is not supported on this path");
        else if (name == "lrn" || name == "localresponsenormalization" || name == "maxout")
            bad(m.first + " belongs to pooling layers, which are outside the hot path");
        else if (name == "sparsenesspenalty") {
            for (auto& q : value.obj) {
                const string k = lower(q.first);
                if (k == "p") nd._sparsenessPenalty_p = (NNFloat)q.second.asNumber();
                else if (k == "beta") nd._sparsenessPenalty_beta = (NNFloat)q.second.asNumber();
                else bad("Invalid SparsenessPenalty parameter: " + q.first);
            }
        } else if (name == "denoising") {
            for (auto& q : value.obj) {
                if (lower(q.first) == "p") nd._denoising_p = (NNFloat)q.second.asNumber();
                else bad("Invalid Denoising parameter: " + q.first);
            }
        } else if (name == "deltaboost") {
            for (auto& q : value.obj) {
                const string k = lower(q.first);
                if (k == "one") nd._deltaBoost_one = (NNFloat)q.second.asNumber();
                else if (k == "zero") nd._deltaBoost_zero = (NNFloat)q.second.asNumber();
                else bad("Invalid DeltaBoost parameter: " + q.first);
            }
        } else if (name == "scaledmarginalcrossentropy" || name == "datascaledmarginalcrossentropy") {
            for (auto& q : value.obj) {
                const string k = lower(q.first);
                if (k == "onetarget") nd._SMCE_oneTarget = (NNFloat)q.second.asNumber();
                else if (k == "zerotarget") nd._SMCE_zeroTarget = (NNFloat)q.second.asNumber();
                else if (k == "onescale") nd._SMCE_oneScale = (NNFloat)q.second.asNumber();
                else if (k == "zeroscale") nd._SMCE_zeroScale = (NNFloat)q.second.asNumber();
                else bad("Invalid ScaledMarginalCrossEntropy parameter: " + q.first);
            }
        } else if (name == "shuffleindices") nd._bShuffleIndices = value.asBool();
        else if (name == "reluslope" || name == "slope") nd._RELUSlope = (NNFloat)value.asNumber();
        else if (name == "elualpha") nd._ELUAlpha = (NNFloat)value.asNumber();
        else if (name == "selulambda") nd._SELULambda = (NNFloat)value.asNumber();
        else if (name == "decay") nd._decay = (NNFloat)value.asNumber();
        else if (name == "errorfunction") {
            const string s = lower(value.asString());
            if (s == "l2") nd._errorFunction = ErrorFunction::L2;
            else if (s == "crossentropy" || s == "cross entropy") nd._errorFunction = ErrorFunction::CrossEntropy;
            else if (s == "scaledmarginalcrossentropy") nd._errorFunction = ErrorFunction::ScaledMarginalCrossEntropy;
            else if (s == "l1" || s == "l2hinge" || s == "hinge" || s == "datascaledmarginalcrossentropy")
                bad("error function " + value.asString() + " is outside the hot path (L2, CrossEntropy, ScaledMarginalCrossEntropy)");
            else bad("Invalid error function: " + value.asString());
        } else if (name == "layers") {
            const size_t size = value.isArray() ? value.arr.size() : 1;
            for (size_t i = 0; i < size; i++) {
                const JValue& layer = value.isArray() ? value.arr[i] : value;
                NNLayerDescriptor ldl;
                bool bSource = false, bAutoSize = false;
                ldl._kind = (i == 0) ? NNLayer::Kind::Input : (i == size - 1 ? NNLayer::Kind::Output : NNLayer::Kind::Hidden);
                ldl._type = NNLayer::Type::FullyConnected;
                for (auto& q : layer.obj) {                                   // kind / type first (E/NNNetwork.cpp:3084-3126)
                    const string k = lower(q.first);
                    if (k == "kind") {
                        const string s = lower(q.second.asString());
                        if (s == "input") ldl._kind = NNLayer::Kind::Input;
                        else if (s == "hidden") ldl._kind = NNLayer::Kind::Hidden;
                        else if (s == "target") ldl._kind = NNLayer::Kind::Target;
                        else if (s == "output") ldl._kind = NNLayer::Kind::Output;
                        else bad("Invalid layer kind: " + q.second.asString());
                    } else if (k == "type") {
                        const string s = lower(q.second.asString());
                        if (s == "fullyconnected") ldl._type = NNLayer::Type::FullyConnected;
                        else if (s == "convolutional" || s == "pooling") bad("layer type " + q.second.asString() + " is outside the hot path");
                        else bad("Invalid layer type: " + q.second.asString());
                    }
                }
                switch (ldl._kind) {                                          // default names, E/NNNetwork.cpp:3141-3158
                case NNLayer::Kind::Input:  ldl._name = "Input" + to_string(nd._vLayerDescriptor.size()); break;
                case NNLayer::Kind::Hidden: ldl._name = "Hidden" + to_string(nd._vLayerDescriptor.size()); break;
                case NNLayer::Kind::Output: ldl._name = "Output" + to_string(nd._vLayerDescriptor.size()); break;
                case NNLayer::Kind::Target: ldl._name = "Target" + to_string(nd._vLayerDescriptor.size()); break;
                }
                for (auto& q : layer.obj) {
                    const string k = lower(q.first);
                    const JValue& lv = q.second;
                    if (k == "kind" || k == "type") continue;
                    if (k == "name") {
                        ldl._name = lv.asString();
                        if (sLayer.count(ldl._name)) bad("Duplicate layer name detected: " + ldl._name);
                        sLayer.insert(ldl._name);
                    } else if (k == "sparse") { if (lv.asBool()) ldl._attributes |= NNLayer::Attributes::Sparse; }
                    else if (k == "n") {
                        if (lv.isArray()) {
                            if (lv.arr.size() > 4 || lv.arr.empty()) bad("N must have 1 to 4 components");
                            ldl._dimensions = (uint32_t)lv.arr.size();
                            if (lv.arr.size() > 3) ldl._Nw = (uint32_t)lv.arr[3].asNumber();
                            if (lv.arr.size() > 2) ldl._Nz = (uint32_t)lv.arr[2].asNumber();
                            if (lv.arr.size() > 1) ldl._Ny = (uint32_t)lv.arr[1].asNumber();
                            ldl._Nx = (uint32_t)lv.arr[0].asNumber();
                        } else if (lv.isString()) {
                            if (lower(lv.asString()) == "auto" && ldl._kind != NNLayer::Kind::Hidden) bAutoSize = true;
                            else if (lower(lv.asString()) == "auto") bad("Illegal attempt to use auto for hidden layer: " + ldl._name);
                            else bad("Invalid N: " + lv.asString());
                        } else { ldl._Nx = (uint32_t)lv.asNumber(); ldl._dimensions = 1; }
                    } else if (k == "pdropout") ldl._pDropout = (NNFloat)lv.asNumber();
                    else if (k == "dataset") ldl._dataSet = lv.asString();
                    else if (k == "source") {
                        if (ldl._kind == NNLayer::Kind::Input) bad("Input layer " + ldl._name + " cannot have a source");
                        if (lv.isArray()) for (auto& e : lv.arr) ldl._vSource.push_back(e.asString());
                        else ldl._vSource.push_back(lv.asString());
                        bSource = true;
                    } else if (k == "activation") ldl._activation = parseActivation(lv.asString());
                    else if (k == "reluslope" || k == "slope") ldl._RELUSlope = (NNFloat)lv.asNumber();
                    else if (k == "elualpha") ldl._ELUAlpha = (NNFloat)lv.asNumber();
                    else if (k == "selulambda") ldl._SELULambda = (NNFloat)lv.asNumber();
                    else if (k == "weightnorm") ldl._weightNorm = (NNFloat)lv.asNumber();
                    else if (k == "deltanorm") ldl._deltaNorm = (NNFloat)lv.asNumber();
                    else if (k == "sparsenesspenalty") {
                        for (auto& z : lv.obj) {
                            const string kk = lower(z.first);
                            if (kk == "p") ldl._sparsenessPenalty_p = (NNFloat)z.second.asNumber();
                            else if (kk == "beta") ldl._sparsenessPenalty_beta = (NNFloat)z.second.asNumber();
                            else bad("Invalid sparseness penalty parameter for layer " + ldl._name + ": " + z.first);
                        }
                    } else if (k == "weightinit") {
                        for (auto& z : lv.obj) {
                            const string kk = lower(z.first);
                            if (kk == "scheme") ldl._weightInit = parseScheme(z.second.asString());
                            else if (kk == "scale") ldl._weightInitScale = (NNFloat)z.second.asNumber();
                            else if (kk == "bias") ldl._biasInit = (NNFloat)z.second.asNumber();
                            else bad("Invalid weight initialization field for layer " + ldl._name + ": " + z.first);
                        }
                    } else if (k == "kernel" || k == "kernelstride" || k == "function" || k == "batchnormalization" || k == "skip" || k == "sharedweights")
                        bad("layer key " + q.first + " (layer " + ldl._name + ") selects a feature outside the hot path");
                    else bad("Unknown neural network layer field: " + q.first);          // unknown key => fatal (E/NNNetwork.cpp:3596)
                }
                if (bAutoSize) {                                               // E/NNNetwork.cpp:3600-3622
                    bool bFound = false;
                    for (auto& p : vDataSet)
                        if (p._name == ldl._dataSet) { ldl._Nx = p._width; ldl._Ny = p._height; ldl._Nz = p._length; ldl._dimensions = p._dimensions; bFound = true; }
                    if (!bFound) bad("Unable to find data set " + ldl._dataSet + " to determine dimensions for layer: " + ldl._name);
                }
                if (!bSource && ldl._kind != NNLayer::Kind::Input) {
                    if (nd._vLayerDescriptor.empty()) bad("layer " + ldl._name + " has no source");
                    ldl._vSource.push_back(nd._vLayerDescriptor.back()._name);
                }
                for (auto& src : ldl._vSource) {
                    NNWeightDescriptor wd;
                    wd._inputLayer = src; wd._outputLayer = ldl._name; wd._norm = ldl._weightNorm;
                    nd._vWeightDescriptor.push_back(wd);
                }
                nd._vLayerDescriptor.push_back(ldl);
            }
        } else bad("Unknown neural network field: " + m.first);                // E/NNNetwork.cpp:3698
    }
    if (nd._sparsenessPenalty_beta > (NNFloat)0.0) nd._bSparsenessPenalty = true;
    if (nd._denoising_p > (NNFloat)0.0) {                                      // E/NNNetwork.cpp:3708-3720
        nd._bDenoising = true;
        for (auto& l : nd._vLayerDescriptor)
            if (l._kind == NNLayer::Kind::Input && (l._attributes & NNLayer::Attributes::Sparse)) l._attributes |= NNLayer::Attributes::Denoising;
    }
    return nd;
}

NNNetwork* LoadNeuralNetworkJSONString(const string& json, const uint32_t batch, const vector<NNDataSetBase*>& vDataSet)
{
    vector<NNDataSetShape> vShape;
    for (auto p : vDataSet) vShape.push_back(NNDataSetShape{p->_name, p->_width, p->_height, p->_length, p->_dimensions});
    NNNetworkDescriptor nd = ParseNeuralNetworkJSON(json, vShape);
    return new NNNetwork(nd, batch);
}

static const char* kindName(NNLayer::Kind k)
{
    switch (k) { case NNLayer::Kind::Input: return "Input"; case NNLayer::Kind::Hidden: return "Hidden"; case NNLayer::Kind::Output: return "Output"; default: return "Target"; }
}
static const char* activationName(Activation a)
{
    static const char* n[] = {"Sigmoid", "Tanh", "RectifiedLinear", "Linear", "ParametricRectifiedLinear", "SoftPlus", "SoftSign", "SoftMax", "RELUMax",
                              "LinearMax", "ExponentialLinear", "LeakyRectifiedLinear", "ScaledExponentialLinear"};
    return n[(int)a];
}
static const char* errorName(ErrorFunction e)
{
    static const char* n[] = {"L1", "L2", "CrossEntropy", "ScaledMarginalCrossEntropy", "DataScaledMarginalCrossEntropy", "Hinge", "L2Hinge"};
    return n[(int)e];
}
static const char* initName(WeightInitialization w)
{
    static const char* n[] = {"Xavier", "CaffeXavier", "Gaussian", "Uniform", "UnitBall", "Constant", "SELU"};
    return n[(int)w];
}

// One line per network / layer / weight, in declaration order -- what the parser understood, for tests and for `train -describe`
string DescribeNeuralNetworkJSON(const string& json, const vector<NNDataSetShape>& vDataSet)
{
    const NNNetworkDescriptor nd = ParseNeuralNetworkJSON(json, vDataSet);
    ostringstream o;
    o << "network name=" << nd._name << " kind=" << (nd._kind == NNNetwork::Kind::AutoEncoder ? "AutoEncoder" : "FeedForward")
      << " error=" << errorName(nd._errorFunction) << " shuffle=" << (nd._bShuffleIndices ? 1 : 0) << " decay=" << nd._decay
      << " denoising_p=" << nd._denoising_p << " sparseness=(" << nd._sparsenessPenalty_p << "," << nd._sparsenessPenalty_beta << ")"
      << " deltaBoost=(" << nd._deltaBoost_one << "," << nd._deltaBoost_zero << ")"
      << " smce=(" << nd._SMCE_oneTarget << "," << nd._SMCE_zeroTarget << "," << nd._SMCE_oneScale << "," << nd._SMCE_zeroScale << ")\n";
    for (auto& l : nd._vLayerDescriptor) {
        o << "layer name=" << l._name << " kind=" << kindName(l._kind) << " N=" << l._Nx << " activation=" << activationName(l._activation)
          << " sparse=" << ((l._attributes & NNLayer::Attributes::Sparse) ? 1 : 0) << " denoising=" << ((l._attributes & NNLayer::Attributes::Denoising) ? 1 : 0)
          << " pDropout=" << l._pDropout << " init=" << initName(l._weightInit) << ":" << l._weightInitScale << ":" << l._biasInit
          << " dataset=" << l._dataSet << " sources=";
        for (size_t i = 0; i < l._vSource.size(); i++) o << (i ? "," : "") << l._vSource[i];
        o << "\n";
    }
    for (auto& w : nd._vWeightDescriptor) o << "weight " << w._inputLayer << " -> " << w._outputLayer << "\n";
    return o.str();
}

NNNetwork* LoadNeuralNetworkJSON(const string& fname, const uint32_t batch, const vector<NNDataSetBase*>& vDataSet)
{
    ifstream f(fname.c_str());
    if (!f.good()) throw DsbEngineError("LoadNeuralNetworkJSON: Failed to open JSON file " + fname);
    stringstream ss;
    ss << f.rdbuf();
    return LoadNeuralNetworkJSONString(ss.str(), batch, vDataSet);
}
