// control.hpp — control plane feeding the hot path from real Sassena inputs ("next" row):
//   scatter.xml  -> Config        reference src/control/parameters.cpp:64-792 (sections sample, stager, scattering, limits, database)
//   db.xml       -> Database      reference src/control/database.cpp:31-145 (parse), :293-528 (evaluation)
//   PDB          -> Atoms         reference src/sample/atoms.cpp:64-86
//   selections   -> index sets    reference src/sample/sample.cpp:30-101, atomselection_reader.cpp:33-118
//   b_j(|q|)     -> ScatterFactors reference src/scatter_devices/scatter_factors.cpp:28-135
// XML is read by a small built-in parser (no libxml2 in this image): elements, text, comments, CDATA, the five
// predefined entities; XPath use is limited to what the reference needs ("//a/b/c", "./x", ".").  XInclude is not
// supported and raises an error.
#pragma once

#include <map>
#include <memory>
#include <regex>
#include <string>
#include <vector>

#include "sassena_host.hpp"

namespace sassena {

// ---- XML --------------------------------------------------------------------------------------------------------
struct XMLNode {
    std::string name;
    std::string text;  // concatenated character data directly inside this element
    std::vector<XMLNode> children;
};

class XMLInterface {  // reference src/io/xml_interface.cpp
    XMLNode root_;
    const XMLNode *current_ = nullptr;

   public:
    explicit XMLInterface(const std::string &filename);
    // base_dir: where relative XInclude hrefs of this document are looked up
    static XMLNode parse(const std::string &text, const std::string &base_dir = std::string());
    std::vector<const XMLNode *> get(const std::string &xpath) const;
    bool exists(const std::string &xpath) const { return !get(xpath).empty(); }
    void set_current(const XMLNode *n) { current_ = n; }
    std::string get_string(const std::string &xpath) const;
    double get_double(const std::string &xpath) const;
    size_t get_size(const std::string &xpath) const;
    long get_long(const std::string &xpath) const;
    bool get_bool(const std::string &xpath) const;  // true/false in any case, or 1/0 (xml_interface.cpp:56-61)
};

// ---- Config (scatter.xml) -----------------------------------------------------------------------------------------
struct SampleSelectionParameters {
    std::string type = "index";  // index | range | lexical | file
    std::string name;
    std::vector<size_t> ids;                // index
    size_t from = 0, to = 0;                // range (inclusive)
    std::string expression;                 // lexical / file
    std::string filepath, format = "pdb", selector = "beta";  // file
};

struct SampleFramesetParameters {
    std::string file = "sample.dcd", filepath, format = "dcd";
    size_t first = 0, last = 0, stride = 1, clones = 1;
    bool last_set = false;
};

// reference structure of an alignment / motion (parameters.cpp:248-281,318-336)
struct SampleReferenceParameters {
    std::string type = "frame";  // frame | file | instant
    size_t frame = 0;
    std::string file, filepath, format = "pdb", selection = "system";
};

struct SampleMotionParameters {  // parameters.cpp:234-296
    std::string type = "linear";  // linear | fixed | oscillation | randomwalk | brownian | rotationalbrownian | localbrownian | none
    double displace = 0.0, frequency = 0.001, radius = 0.0;
    CartesianCoor3D direction{1, 0, 0};
    std::string selection = "system";
    unsigned long seed = 0;
    long sampling = 1;
    SampleReferenceParameters reference;  // default type "instant"
};

struct SampleAlignmentParameters {  // parameters.cpp:298-341
    std::string type = "center";  // center | fittrans | fitrottrans | fitrot
    std::string selection = "system", order = "pre";
    SampleReferenceParameters reference;  // default type "frame", frame 0
};

struct ScatteringBackgroundKappaParameters {
    std::string selection = "system";
    double value = 1.0;
};

struct Config : Params {
    std::string config_rootpath;
    // sample
    std::string structure_file = "sample.pdb", structure_filepath, structure_format = "pdb";
    std::vector<SampleSelectionParameters> selections;
    std::vector<SampleFramesetParameters> framesets;
    std::vector<SampleMotionParameters> motions;
    std::vector<SampleAlignmentParameters> alignments;
    // stager
    std::string stager_target = "system";
    // scattering
    std::vector<CartesianCoor3D> qvectors;
    std::vector<ScatteringVectorsScan> scans;
    double background_factor = 0.0;
    std::vector<ScatteringBackgroundKappaParameters> kappas;
    std::string signal_file = "signal.h5", signal_filepath;
    bool signal_fqt = true, signal_fq0 = true, signal_fq = true, signal_fq2 = true;
    size_t signal_chunksize = 10000;  // limits.signal.chunksize (parameters.cpp:626)
    size_t signal_flush_seconds = 600;  // limits.services.signal.times.serverflush (parameters.cpp:628,703-705)
    std::string rawconfig;            // the configuration file as read (Params::get_rawconfig, meta/rawconfig + meta/config)
    // database
    std::string database_file = "db.xml", database_filepath, database_format = "xml";

    std::string get_filepath(const std::string &fname) const;  // parameters.cpp:41-54
    void read_xml(const std::string &filename);                // parameters.cpp:64-792
};

// ---- Database (db.xml) --------------------------------------------------------------------------------------------
class Database {
    std::map<std::string, std::string> label2regexp_;  // names/pdb
    std::map<std::string, size_t> ids_;
    std::map<size_t, std::string> rids_;
    size_t next_id_ = 0;
    std::map<size_t, double> masses_;
    struct Fn {
        std::vector<double> v;
        size_t type = 0;
    };
    std::map<size_t, Fn> volumes_, exclusion_, sfactors_;
    std::map<std::string, std::string> quick_;

   public:
    std::string rawconfig;  // the database file as read (Database::get_config, meta/database)
    void read_xml(const std::string &filename);
    std::string pdb_name(const std::string &testlabel);  // database.cpp:308-340: exactly one regex must match
    size_t atom_id(const std::string &label);            // atomIDs.get (registers unknown labels)
    std::string atom_label(size_t id) const;
    double mass(size_t id) const;
    double volume(size_t id) const;                                   // database.cpp:391-411
    double exclusionfactor(size_t id, double effvolume, double q) const;  // :430-450
    double sfactor(size_t id, double q) const;                        // :469-528
};

// ---- Sample ---------------------------------------------------------------------------------------------------------
struct LoadedSample {
    std::vector<size_t> atom_ids;  // database element ID per atom (Atoms::ids_)
    std::map<std::string, std::vector<size_t>> selections;
    size_t NF = 0;
    std::vector<float> frames;  // [NF][NA_target][3], reduced to stager.target (CoordinateSets::load, coordinate_sets.cpp:353-357)
    std::vector<size_t> target;  // atom indices of stager.target
};

// Atoms::add (atoms.cpp:64-86): "ATOM  " records, name = columns 13-16 trimmed -> database label -> ID
std::vector<size_t> read_pdb_atoms(const std::string &filename, Database &db);
// Sample::init (sample.cpp:30-101): selections + the reserved "system" selection
void init_selections(const Config &cfg, Database &db, LoadedSample &s, const std::string &structure_path);
// CoordinateSets: framesets (dcd / dcdlist / pdb / pdblist / xtc / trr with first/last/stride/clones), every frame passed
// through the alignments and motions of the sample section (CoordinateSets::load, coordinate_sets.cpp:245-368), then
// reduced to the target selection and narrowed to float (data_stager.cpp:111-113)
void load_frames(const Config &cfg, const Database &db, LoadedSample &s);

// Motion walkers (reference src/sample/motion_walker.cpp): the 4x4 transform of frame `timepos`, applied to row vectors
// (x, y, z, 1) * T (CartesianCoordinateSet::transform, coordinate_set.cpp:178-215).
class MotionWalker {
   public:
    virtual ~MotionWalker() {}
    virtual void transform(size_t timepos, double T[16]) = 0;
    static MotionWalker *create(const SampleMotionParameters &m);  // nullptr for type "none"; throws for unknown types
};

// CoordinateSets::load between reading a frame and reducing it to the target (coordinate_sets.cpp:245-353): pre-alignments,
// motions, post-alignments, all in double on the whole system.
class CoordinateSetsProcessor {
   public:
    struct Alignment {
        std::string type, selection, reference_selection;
        bool has_reference = false;
        std::vector<double> ref;  // [n][3] coordinates of the reference selection
    };
    struct Motion {
        std::string selection, reference_selection;
        bool has_reference = false;
        std::vector<double> ref;
        std::unique_ptr<MotionWalker> walker;
    };
    // load_raw(framenumber, xyz[natoms*3]) reads the unprocessed frame (reference type "frame")
    CoordinateSetsProcessor(const Config &cfg, const Database &db, const LoadedSample &s, size_t NF,
                            const std::function<void(size_t, float *)> &load_raw);
    bool active() const { return !pre_.empty() || !post_.empty() || !motions_.empty(); }
    void apply(size_t framenumber, std::vector<double> &xyz) const;  // xyz[natoms*3] in place

   private:
    const Database &db_;
    const LoadedSample &s_;
    std::vector<Alignment> pre_, post_;
    std::vector<Motion> motions_;
    const std::vector<size_t> &sel(const std::string &name) const;
    std::vector<double> reference_set(const SampleReferenceParameters &r, const Config &cfg, size_t NF,
                                      const std::function<void(size_t, float *)> &load_raw) const;
    void align(const Alignment &a, std::vector<double> &xyz) const;
};

// ScatterFactors (scatter_factors.cpp:28-135): b_j(|q|) = sf - background.factor * excl(ID, kappa*V, |q|)
class ScatterFactors {
    const Database &db_;
    const LoadedSample &sample_;
    std::vector<double> kappas_;
    double background_;

   public:
    ScatterFactors(const Config &cfg, const Database &db, const LoadedSample &s);
    void update(double ql, double *factors) const;       // one entry per atom of the target selection
    double compute_background(double ql) const;          // scatter_factors.cpp:104-124
};

// .npy writer for the interim signal output (datasets named like the HDF5 layout, file_writer_service.cpp:44-171)
void write_npy(const std::string &path, const double *data, const std::vector<size_t> &shape);
bool read_npy(const std::string &path, std::vector<double> &data, std::vector<size_t> &shape);

// the `sassena` executable's flow (src/main/sassena.cpp:132-417) for one process / one communicator:
// load(): config + database + sample;  run(): factory -> device.run() -> signal directory (.npy datasets).
struct Job {
    Config cfg;
    Database db;
    LoadedSample sample;
    std::unique_ptr<ScatterFactors> factors;
    // `overwrites`: the reference's command-line overwrite options (Params::options / overwrite_options,
    // parameters.cpp:795-875), (key, value) pairs applied after the configuration file has been read and before the
    // database and the sample are loaded.  Keys: sample.structure.file, sample.structure.format, stager.target,
    // stager.dump, stager.file, stager.format, scattering.signal.file, limits.computation.threads.
    void load(const std::string &config_file, const std::vector<std::pair<std::string, std::string>> &overwrites = {});
    // returns the number of q-vectors this rank wrote; with more than one rank every writing rank stores its rows
    // under <signal_dir>/rank_<r>/.  A `signal_dir` ending in ".h5" selects the reference's output: the rows go to
    // "<path>.d/" as above and rank 0 then writes <path> as an HDF5 file in the layout of
    // file_writer_service.cpp:44-171; if <path> already exists its q-vectors are skipped and its rows kept (resume,
    // sassena.cpp:270-305).
    size_t run(const std::string &signal_dir, std::shared_ptr<ICommunicator> comm, const SgpuBackend &be, sgpu_ctx *ctx,
               std::string *report);
    // the `s_stage` executable's flow (src/main/s_stage.cpp:205-232): stage the trajectory of stager.target on the ranks of
    // `comm` the way stager.mode says ("frames": DataStagerByFrame, every rank its DivAssignment block of frames; "atoms":
    // DataStagerByAtom, every rank its ModAssignment atoms) and, with stager.dump, write the staged coordinates to
    // stager.file.  Returns the bytes this rank staged.
    size_t stage(std::shared_ptr<ICommunicator> comm, const SgpuBackend &be, sgpu_ctx *ctx, std::string *report);
};

}  // namespace sassena
