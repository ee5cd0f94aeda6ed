// control.cpp — see control.hpp.  Reference citations are relative to benlabs/sassena v1.4.2.
#include "control.hpp"

#include <dirent.h>
#include <sys/stat.h>
#include <unistd.h>

#include <chrono>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>

#include "dcd.hpp"
#include "h5mini.hpp"
#include "xdr_traj.hpp"

namespace sassena {

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

namespace {
std::string trim(const std::string &s) {
    size_t a = 0, b = s.size();
    while (a < b && isspace((unsigned char)s[a])) a++;
    while (b > a && isspace((unsigned char)s[b - 1])) b--;
    return s.substr(a, b - a);
}

std::string read_file(const std::string &fn) {
    std::ifstream f(fn.c_str(), std::ios::binary);
    if (!f) throw Error("cannot open file: " + fn);
    std::stringstream ss;
    ss << f.rdbuf();
    return ss.str();
}

std::string decode_entities(const std::string &s) {
    std::string o;
    for (size_t i = 0; i < s.size(); i++) {
        if (s[i] == '&') {
            size_t e = s.find(';', i);
            if (e != std::string::npos) {
                std::string ent = s.substr(i + 1, e - i - 1);
                const char *r = nullptr;
                if (ent == "lt") r = "<";
                else if (ent == "gt") r = ">";
                else if (ent == "amp") r = "&";
                else if (ent == "quot") r = "\"";
                else if (ent == "apos") r = "'";
                if (r) {
                    o += r;
                    i = e;
                    continue;
                }
            }
        }
        o += s[i];
    }
    return o;
}

// ---- recursive-descent XML parser ----
struct Parser {
    const std::string &t;
    size_t p = 0;
    explicit Parser(const std::string &text) : t(text) {}
    [[noreturn]] void fail(const std::string &m) const { throw Error("XML parse error at byte " + std::to_string(p) + ": " + m); }
    bool starts(const char *s) const { return t.compare(p, strlen(s), s) == 0; }
    void skip_ws() {
        while (p < t.size() && isspace((unsigned char)t[p])) p++;
    }
    void skip_until(const char *end) {
        size_t e = t.find(end, p);
        if (e == std::string::npos) fail(std::string("unterminated construct, expected ") + end);
        p = e + strlen(end);
    }
    void skip_misc() {  // prolog / comments / processing instructions / doctype
        for (;;) {
            skip_ws();
            if (starts("<?")) skip_until("?>");
            else if (starts("<!--")) skip_until("-->");
            else if (starts("<!DOCTYPE")) skip_until(">");
            else break;
        }
    }
    std::string name() {
        size_t s = p;
        while (p < t.size() && (isalnum((unsigned char)t[p]) || t[p] == '_' || t[p] == '-' || t[p] == ':' || t[p] == '.')) p++;
        if (p == s) fail("element name expected");
        return t.substr(s, p - s);
    }
    // XInclude (reference src/io/xml_interface.cpp:109-116 runs libxml2's xmlXIncludeProcess on every document): an
    // <xi:include href="..." [parse="xml"|"text"]/> element is replaced by the root element of the referenced document
    // (href relative to the including file) or by its text; a missing document takes the element's <xi:fallback> content.
    // xpointer selections are not supported.
    std::string base_dir;  // directory of the document being parsed ("" = current directory)
    int depth = 0;         // inclusion nesting
    typedef std::vector<std::pair<std::string, std::string>> Attrs;
    static std::string local_name(const std::string &n) {
        size_t c = n.find(':');
        return c == std::string::npos ? n : n.substr(c + 1);
    }
    static const std::string *attr(const Attrs &a, const char *key) {
        for (auto &kv : a)
            if (kv.first == key) return &kv.second;
        return nullptr;
    }
    // appends what the include element stands for to `parent`
    void include_into(XMLNode &parent, const Attrs &a, const XMLNode &self) {
        if (attr(a, "xpointer")) throw Error("XInclude: xpointer is not supported by the built-in XML reader");
        const std::string *href = attr(a, "href");
        if (!href || href->empty()) throw Error("XInclude: include element without href");
        const std::string *mode = attr(a, "parse");
        if (depth >= 16) throw Error("XInclude: inclusion nested deeper than 16 levels (a loop?): " + *href);
        std::string fn = *href;
        if (fn.compare(0, 7, "file://") == 0) fn = fn.substr(7);
        if (!fn.empty() && fn[0] != '/' && !base_dir.empty()) fn = base_dir + "/" + fn;
        std::string text;
        {
            std::ifstream f(fn.c_str(), std::ios::binary);
            if (!f) {
                for (auto &c : self.children)
                    if (local_name(c.name) == "fallback") {
                        parent.text += c.text;
                        for (auto &g : c.children) parent.children.push_back(g);
                        return;
                    }
                throw Error("XInclude: cannot open file: " + fn);
            }
            std::stringstream ss;
            ss << f.rdbuf();
            text = ss.str();
        }
        if (mode && *mode == "text") {
            parent.text += text;
            return;
        }
        if (mode && *mode != "xml") throw Error("XInclude: unknown parse mode: " + *mode);
        Parser sub(text);
        size_t slash = fn.rfind('/');
        sub.base_dir = slash == std::string::npos ? std::string() : fn.substr(0, slash);
        sub.depth = depth + 1;
        sub.skip_misc();
        if (sub.p >= text.size()) throw Error("XInclude: no root element in " + fn);
        parent.children.push_back(sub.element());
    }
    XMLNode element(Attrs *attrs_out = nullptr) {
        if (t[p] != '<') fail("'<' expected");
        p++;
        XMLNode n;
        n.name = name();
        // attributes are skipped (the reference's schema carries everything in elements) unless the caller asks for them
        for (;;) {
            skip_ws();
            if (p >= t.size()) fail("unterminated start tag");
            if (t[p] == '/' && p + 1 < t.size() && t[p + 1] == '>') {
                p += 2;
                return n;
            }
            if (t[p] == '>') {
                p++;
                break;
            }
            const std::string key = name();
            skip_ws();
            if (p < t.size() && t[p] == '=') {
                p++;
                skip_ws();
                if (p >= t.size() || (t[p] != '"' && t[p] != '\'')) fail("quoted attribute value expected");
                char q = t[p++];
                size_t e = t.find(q, p);
                if (e == std::string::npos) fail("unterminated attribute value");
                if (attrs_out) attrs_out->push_back(std::make_pair(key, decode_entities(t.substr(p, e - p))));
                p = e + 1;
            }
        }
        for (;;) {
            if (p >= t.size()) fail("unterminated element <" + n.name + ">");
            if (starts("</")) {
                p += 2;
                std::string closing = name();
                if (closing != n.name) fail("mismatched closing tag </" + closing + "> for <" + n.name + ">");
                skip_ws();
                if (p >= t.size() || t[p] != '>') fail("'>' expected");
                p++;
                return n;
            }
            if (starts("<!--")) skip_until("-->");
            else if (starts("<![CDATA[")) {
                size_t e = t.find("]]>", p);
                if (e == std::string::npos) fail("unterminated CDATA");
                n.text += t.substr(p + 9, e - p - 9);
                p = e + 3;
            } else if (starts("<?")) skip_until("?>");
            else if (t[p] == '<') {
                Attrs a;
                XMLNode c = element(&a);
                if (local_name(c.name) == "include" && c.name != "include") include_into(n, a, c);  // prefixed: xi:include
                else n.children.push_back(c);
            } else {
                size_t e = t.find('<', p);
                if (e == std::string::npos) fail("unterminated element <" + n.name + ">");
                n.text += decode_entities(t.substr(p, e - p));
                p = e;
            }
        }
    }
};

void find_all(const XMLNode &n, const std::string &name, std::vector<const XMLNode *> &out) {
    if (n.name == name) out.push_back(&n);
    for (auto &c : n.children) find_all(c, name, out);
}

std::vector<std::string> split_path(const std::string &s) {
    std::vector<std::string> parts;
    std::string cur;
    for (char c : s) {
        if (c == '/') {
            if (!cur.empty()) parts.push_back(cur);
            cur.clear();
        } else cur += c;
    }
    if (!cur.empty()) parts.push_back(cur);
    return parts;
}
}  // namespace

XMLNode XMLInterface::parse(const std::string &text, const std::string &base_dir) {
    Parser ps(text);
    ps.base_dir = base_dir;
    ps.skip_misc();
    if (ps.p >= text.size()) throw Error("XML parse error: no root element");
    return ps.element();
}

XMLInterface::XMLInterface(const std::string &filename)
    : root_(parse(read_file(filename), filename.rfind('/') == std::string::npos ? std::string() : filename.substr(0, filename.rfind('/')))) {}

std::vector<const XMLNode *> XMLInterface::get(const std::string &xpath) const {
    std::vector<const XMLNode *> cur;
    std::vector<std::string> parts;
    if (xpath.compare(0, 2, "//") == 0) {
        parts = split_path(xpath.substr(2));
        if (parts.empty()) return cur;
        find_all(root_, parts[0], cur);
        parts.erase(parts.begin());
    } else if (xpath == ".") {
        if (current_) cur.push_back(current_);
        return cur;
    } else if (xpath.compare(0, 2, "./") == 0) {
        if (!current_) throw Error("XML: relative path without a current node: " + xpath);
        cur.push_back(current_);
        parts = split_path(xpath.substr(2));
    } else {
        throw Error("XML: unsupported path expression: " + xpath);
    }
    for (auto &part : parts) {
        std::vector<const XMLNode *> next;
        for (auto *n : cur)
            for (auto &c : n->children)
                if (c.name == part) next.push_back(&c);
        cur = next;
    }
    return cur;
}

std::string XMLInterface::get_string(const std::string &xpath) const {
    auto v = get(xpath);
    if (v.empty()) throw Error("XML: path does not resolve: " + xpath);
    return trim(v[0]->text);
}
double XMLInterface::get_double(const std::string &xpath) const {
    const std::string s = get_string(xpath);
    char *e = nullptr;
    double v = strtod(s.c_str(), &e);
    if (e == s.c_str()) throw Error("XML: bad numeric value '" + s + "' at " + xpath);
    return v;
}
size_t XMLInterface::get_size(const std::string &xpath) const {
    const std::string s = get_string(xpath);
    char *e = nullptr;
    unsigned long long v = strtoull(s.c_str(), &e, 10);
    if (e == s.c_str()) throw Error("XML: bad integer value '" + s + "' at " + xpath);
    return (size_t)v;
}
long XMLInterface::get_long(const std::string &xpath) const {
    const std::string s = get_string(xpath);
    char *e = nullptr;
    long v = strtol(s.c_str(), &e, 10);
    if (e == s.c_str()) throw Error("XML: bad integer value '" + s + "' at " + xpath);
    return v;
}
bool XMLInterface::get_bool(const std::string &xpath) const {
    std::string s = get_string(xpath), u;
    for (char c : s) u += (char)toupper(c);
    if (u == "TRUE" || u == "1") return true;
    if (u == "FALSE" || u == "0") return false;
    throw Error("XML: bad boolean value '" + s + "' at " + xpath);
}

// ---------------------------------------------------------------------------------------------------------------
// Config
// ---------------------------------------------------------------------------------------------------------------
std::string Config::get_filepath(const std::string &fname) const {
    if (!fname.empty() && fname[0] == '/') return fname;
    if (config_rootpath.empty()) return fname;
    return config_rootpath + "/" + fname;
}

void Config::read_xml(const std::string &filename) {
    {
        size_t slash = filename.find_last_of('/');
        config_rootpath = (slash == std::string::npos) ? std::string(".") : filename.substr(0, slash);
    }
    {
        std::ifstream in(filename.c_str(), std::ios::binary);
        std::stringstream ss;
        ss << in.rdbuf();
        rawconfig = ss.str();
    }
    XMLInterface x(filename);
    // ---- sample (parameters.cpp:85-343) ----
    structure_filepath = get_filepath(structure_file);
    if (x.exists("//sample/structure/file")) {
        structure_file = x.get_string("//sample/structure/file");
        structure_filepath = get_filepath(structure_file);
    }
    if (x.exists("//sample/structure/format")) structure_format = x.get_string("//sample/structure/format");
    {
        auto sels = x.get("//sample/selections/selection");
        for (size_t i = 0; i < sels.size(); ++i) {
            x.set_current(sels[i]);
            SampleSelectionParameters sp;
            if (x.exists("./type")) sp.type = x.get_string("./type");
            sp.name = "_" + std::to_string(i);
            if (x.exists("./name")) sp.name = x.get_string("./name");
            if (sp.type == "index") {
                for (auto *n : x.get("./index")) {
                    x.set_current(n);
                    sp.ids.push_back(x.get_size("."));
                    x.set_current(sels[i]);
                }
            } else if (sp.type == "range") {
                if (x.exists("./from")) sp.from = x.get_size("./from");
                if (x.exists("./to")) sp.to = x.get_size("./to");
            } else if (sp.type == "lexical") {
                if (x.exists("./expression")) sp.expression = x.get_string("./expression");
            } else if (sp.type == "file") {
                std::string file = "selection.pdb";
                sp.expression = "1|1\\.0|1\\.00";
                if (x.exists("./format")) sp.format = x.get_string("./format");
                if (sp.format == "ndx") {
                    file = "index.ndx";
                    sp.selector = "name";
                    sp.expression = ".*";
                }
                if (x.exists("./file")) file = x.get_string("./file");
                sp.filepath = get_filepath(file);
                if (x.exists("./selector")) sp.selector = x.get_string("./selector");
                if (x.exists("./expression")) sp.expression = x.get_string("./expression");
            } else {
                throw Error("Selection type not understood: " + sp.type);
            }
            selections.push_back(sp);
        }
    }
    if (x.exists("//sample/framesets")) {
        SampleFramesetParameters def;
        if (x.exists("//sample/framesets/first")) def.first = x.get_size("//sample/framesets/first");
        if (x.exists("//sample/framesets/last")) {
            def.last = x.get_size("//sample/framesets/last");
            def.last_set = true;
        }
        if (x.exists("//sample/framesets/stride")) def.stride = x.get_size("//sample/framesets/stride");
        if (x.exists("//sample/framesets/clones")) def.clones = x.get_size("//sample/framesets/clones");
        // (the reference parses framesets/format into def_first, parameters.cpp:192; the default format stays "dcd")
        for (auto *n : x.get("//sample/framesets/frameset")) {
            x.set_current(n);
            SampleFramesetParameters f = def;
            if (x.exists("./file")) f.file = x.get_string("./file");
            f.filepath = get_filepath(f.file);
            if (x.exists("./format")) f.format = x.get_string("./format");
            if (x.exists("./first")) f.first = x.get_size("./first");
            if (x.exists("./last")) {
                f.last = x.get_size("./last");
                f.last_set = true;
            }
            if (x.exists("./stride")) f.stride = x.get_size("./stride");
            if (x.exists("./clones")) f.clones = x.get_size("./clones");
            framesets.push_back(f);
        }
    }
    auto read_reference = [&](SampleReferenceParameters &r) {  // parameters.cpp:269-280,325-335
        if (!x.exists("./reference")) return;
        if (x.exists("./reference/type")) r.type = x.get_string("./reference/type");
        if (x.exists("./reference/frame")) r.frame = x.get_size("./reference/frame");
        if (x.exists("./reference/file")) {
            r.file = x.get_string("./reference/file");
            r.filepath = get_filepath(r.file);
        }
        if (x.exists("./reference/format")) r.format = x.get_string("./reference/format");
        if (x.exists("./reference/selection")) r.selection = x.get_string("./reference/selection");
    };
    for (auto *n : x.get("//sample/motions/motion")) {  // parameters.cpp:234-296
        x.set_current(n);
        SampleMotionParameters m;
        m.reference.type = "instant";
        m.reference.file = structure_file;
        m.reference.filepath = get_filepath(structure_file);
        m.reference.format = structure_format;
        if (x.exists("./type")) m.type = x.get_string("./type");
        if (x.exists("./displace")) m.displace = x.get_double("./displace");
        if (x.exists("./frequency")) m.frequency = x.get_double("./frequency");
        if (x.exists("./seed")) m.seed = (unsigned long)x.get_size("./seed");
        if (x.exists("./sampling")) m.sampling = x.get_long("./sampling");
        if (x.exists("./selection")) m.selection = x.get_string("./selection");
        if (x.exists("./direction/x")) m.direction.x = x.get_double("./direction/x");
        if (x.exists("./direction/y")) m.direction.y = x.get_double("./direction/y");
        if (x.exists("./direction/z")) m.direction.z = x.get_double("./direction/z");
        m.radius = m.displace * 10;
        if (x.exists("./radius")) m.radius = x.get_double("./radius");
        // (the reference takes the selection BEFORE ./selection is read as the default reference selection, :253: "system")
        m.reference.selection = "system";
        read_reference(m.reference);
        motions.push_back(m);
    }
    for (auto *n : x.get("//sample/alignments/alignment")) {  // parameters.cpp:298-341
        x.set_current(n);
        SampleAlignmentParameters a;
        if (x.exists("./type")) a.type = x.get_string("./type");
        if (x.exists("./selection")) a.selection = x.get_string("./selection");
        if (x.exists("./order")) a.order = x.get_string("./order");
        a.reference.type = "frame";
        a.reference.frame = 0;
        a.reference.file = structure_file;
        a.reference.filepath = get_filepath(structure_file);
        a.reference.format = structure_format;
        a.reference.selection = a.selection;
        read_reference(a.reference);
        alignments.push_back(a);
    }
    // ---- stager (parameters.cpp:345-369) ----
    if (x.exists("//stager/target")) stager_target = x.get_string("//stager/target");
    stager.filepath = get_filepath(stager.file);
    if (x.exists("//stager/dump")) stager.dump = x.get_bool("//stager/dump");
    if (x.exists("//stager/file")) {
        stager.file = x.get_string("//stager/file");
        stager.filepath = get_filepath(stager.file);
    }
    if (x.exists("//stager/format")) stager.format = x.get_string("//stager/format");
    if (x.exists("//stager/mode")) stager.mode = x.get_string("//stager/mode");  // (the devices force their own mode)
    // ---- scattering (parameters.cpp:372-606) ----
    if (x.exists("//scattering/type")) scattering.type = x.get_string("//scattering/type");
    if (x.exists("//scattering/target")) throw Error("scattering.target is obsolete. Use stager.target instead.");
    if (x.exists("//scattering/background/factor")) background_factor = x.get_double("//scattering/background/factor");
    for (auto *n : x.get("//scattering/background/kappas/kappa")) {
        x.set_current(n);
        ScatteringBackgroundKappaParameters k;
        if (x.exists("./selection")) k.selection = x.get_string("./selection");
        if (x.exists("./value")) k.value = x.get_double("./value");
        kappas.push_back(k);
    }
    if (x.exists("//scattering/vectors")) {
        std::string vt = "single";
        if (x.exists("//scattering/vectors/type")) vt = x.get_string("//scattering/vectors/type");
        if (vt == "single") {
            double qx = 1.0, qy = 0, qz = 0;
            if (x.exists("//scattering/vectors/single/x")) {
                qx = x.get_double("//scattering/vectors/single/x");
                if (x.exists("//scattering/vectors/single/y")) qy = x.get_double("//scattering/vectors/single/y");
                if (x.exists("//scattering/vectors/single/z")) qz = x.get_double("//scattering/vectors/single/z");
            }
            qvectors.push_back(CartesianCoor3D(qx, qy, qz));
        } else if (vt == "scans") {
            for (auto *n : x.get("//scattering/vectors/scans/scan")) {
                x.set_current(n);
                ScatteringVectorsScan sc;
                if (x.exists("./base/x")) sc.basevector.x = x.get_double("./base/x");
                if (x.exists("./base/y")) sc.basevector.y = x.get_double("./base/y");
                if (x.exists("./base/z")) sc.basevector.z = x.get_double("./base/z");
                if (x.exists("./from")) sc.from = x.get_double("./from");
                if (x.exists("./to")) sc.to = x.get_double("./to");
                if (x.exists("./points")) sc.points = x.get_size("./points");
                if (x.exists("./exponent")) sc.exponent = x.get_double("./exponent");
                scans.push_back(sc);
            }
            qvectors = create_from_scans(scans);
        } else if (vt == "file") {
            std::string qf = "qvectors.txt";
            if (x.exists("//scattering/vectors/file")) qf = x.get_string("//scattering/vectors/file");
            std::ifstream in(get_filepath(qf).c_str());
            double a, b, c;
            while (in >> a >> b >> c) qvectors.push_back(CartesianCoor3D(a, b, c));
        }
    }
    if (qvectors.empty()) throw Error("No q vectors generated. Check the scattering.vectors section for errors.");
    if (x.exists("//scattering/dsp/type")) scattering.dsp_type = x.get_string("//scattering/dsp/type");
    if (x.exists("//scattering/dsp/method")) scattering.dsp_method = x.get_string("//scattering/dsp/method");
    const std::string o = "//scattering/average/orientation/";
    if (x.exists(o + "axis")) {
        scattering.axis.x = x.get_double(o + "axis/x");
        scattering.axis.y = x.get_double(o + "axis/y");
        scattering.axis.z = x.get_double(o + "axis/z");
    }
    if (x.exists(o + "type")) scattering.orientation_type = x.get_string(o + "type");
    if (scattering.orientation_type == "vectors") {
        OrientationVectorsParameters &v = scattering.vectors;
        if (x.exists(o + "vectors/type")) v.type = x.get_string(o + "vectors/type");
        if (x.exists(o + "vectors/algorithm")) v.algorithm = x.get_string(o + "vectors/algorithm");
        if (x.exists(o + "vectors/resolution")) v.resolution = (size_t)x.get_long(o + "vectors/resolution");
        if (x.exists(o + "vectors/seed")) v.seed = (uint32_t)x.get_size(o + "vectors/seed");
        if (v.type == "file") {
            std::string vf = "qvector-orientations.txt";
            if (x.exists(o + "vectors/file")) vf = x.get_string(o + "vectors/file");
            std::ifstream in(get_filepath(vf).c_str());
            double a, b, c;
            while (in >> a >> b >> c) v.vectors.push_back(CartesianCoor3D(a, b, c));
        }
        v.create();
    } else if (scattering.orientation_type == "multipole") {
        OrientationMultipoleParameters &m = scattering.multipole;
        if (x.exists(o + "multipole/type")) m.type = x.get_string(o + "multipole/type");
        if (x.exists(o + "multipole/moments/type")) m.moments_type = x.get_string(o + "multipole/moments/type");
        if (x.exists(o + "multipole/moments/resolution")) m.resolution = x.get_long(o + "multipole/moments/resolution");
        if (m.moments_type == "file") {
            std::string mf = "moments.txt";
            if (x.exists(o + "multipole/moments/file")) mf = x.get_string(o + "multipole/moments/file");
            std::ifstream in(get_filepath(mf).c_str());
            long a, b;
            while (in >> a >> b) m.moments.push_back(std::make_pair(a, b));
        }
        m.create();
    } else if (scattering.orientation_type != "none") {
        throw Error("Orientation averaging type not understood: " + scattering.orientation_type);
    }
    signal_filepath = get_filepath(signal_file);
    if (x.exists("//scattering/signal/file")) {
        signal_file = x.get_string("//scattering/signal/file");
        signal_filepath = get_filepath(signal_file);
    }
    if (x.exists("//scattering/signal/fqt")) signal_fqt = x.get_bool("//scattering/signal/fqt");
    if (x.exists("//scattering/signal/fq0")) signal_fq0 = x.get_bool("//scattering/signal/fq0");
    if (x.exists("//scattering/signal/fq")) signal_fq = x.get_bool("//scattering/signal/fq");
    if (x.exists("//scattering/signal/fq2")) signal_fq2 = x.get_bool("//scattering/signal/fq2");
    // ---- limits (parameters.cpp:612-736): the keys the GPU path keeps ----
    if (x.exists("//limits/stage/memory/data")) limits.stage_memory_data = x.get_size("//limits/stage/memory/data");
    if (x.exists("//limits/signal/chunksize")) signal_chunksize = x.get_size("//limits/signal/chunksize");
    if (x.exists("//limits/services/signal/times/serverflush"))
        signal_flush_seconds = x.get_size("//limits/services/signal/times/serverflush");
    if (x.exists("//limits/stage/stream")) limits.stage_stream = x.get_bool("//limits/stage/stream");
    if (x.exists("//limits/decomposition/utilization"))
        limits.decomposition.utilization = x.get_double("//limits/decomposition/utilization");
    if (x.exists("//limits/decomposition/partitions/automatic"))
        limits.decomposition.partitions_automatic = x.get_bool("//limits/decomposition/partitions/automatic");
    if (x.exists("//limits/computation/scan")) limits.coherent_scan = x.get_size("//limits/computation/scan");
    if (x.exists("//limits/computation/scan_snap")) limits.coherent_scan_snap = x.get_bool("//limits/computation/scan_snap");
    if (x.exists("//limits/decomposition/coherent")) limits.coherent_sharding = x.get_string("//limits/decomposition/coherent");
    if (x.exists("//limits/decomposition/partitions/size"))
        limits.decomposition.partitions_size = x.get_size("//limits/decomposition/partitions/size");
    // ---- database (parameters.cpp:776-791) ----
    database_filepath = get_filepath(database_file);
    if (x.exists("//database/file")) {
        database_file = x.get_string("//database/file");
        database_filepath = get_filepath(database_file);
    }
    if (x.exists("//database/format")) database_format = x.get_string("//database/format");
}

// ---------------------------------------------------------------------------------------------------------------
// Database
// ---------------------------------------------------------------------------------------------------------------
size_t Database::atom_id(const std::string &label) {
    auto it = ids_.find(label);
    if (it != ids_.end()) return it->second;
    ids_[label] = next_id_;
    rids_[next_id_] = label;
    return next_id_++;
}
std::string Database::atom_label(size_t id) const {
    auto it = rids_.find(id);
    return it == rids_.end() ? std::string("<Unknown>") : it->second;
}

void Database::read_xml(const std::string &filename) {
    {
        std::ifstream in(filename.c_str(), std::ios::binary);
        std::stringstream ss;
        ss << in.rdbuf();
        rawconfig = ss.str();
    }
    XMLInterface x(filename);
    auto params = [&](std::vector<double> &values) {
        const XMLNode *el = x.get(".")[0];
        for (auto *pn : x.get("./param")) {
            x.set_current(pn);
            values.push_back(x.get_double("."));
            x.set_current(el);
        }
    };
    for (auto *n : x.get("//names/pdb/element")) {
        x.set_current(n);
        label2regexp_[x.get_string("./name")] = x.get_string("./param");
    }
    for (auto *n : x.get("//masses/element")) {
        x.set_current(n);
        masses_[atom_id(x.get_string("./name"))] = x.get_double("./param");
    }
    auto read_fn = [&](const char *path, std::map<size_t, Fn> &dst) {
        for (auto *n : x.get(path)) {
            x.set_current(n);
            Fn f;
            f.type = x.get_size("./type");
            params(f.v);
            dst[atom_id(x.get_string("./name"))] = f;
        }
    };
    read_fn("//sizes/element", volumes_);
    read_fn("//exclusionfactors/element", exclusion_);
    if (!x.exists("//scatterfactors")) throw Error("Need scattering factors.");
    read_fn("//scatterfactors/element", sfactors_);
}

std::string Database::pdb_name(const std::string &testlabel) {
    auto q = quick_.find(testlabel);
    if (q != quick_.end()) return q->second;
    int matches = 0;
    std::string label = "<Unknown>";
    for (auto &lr : label2regexp_) {
        std::regex e(lr.second);
        if (std::regex_match(testlabel, e)) {
            if (matches > 0)
                throw Error("PDB atom name matches" + label + " and " + lr.first + " expression to match: " + testlabel);
            label = lr.first;
            matches++;
        }
    }
    if (matches == 0) throw Error("PDB atom name not recognized: " + testlabel);
    quick_[testlabel] = label;
    return label;
}

double Database::mass(size_t id) const {
    auto it = masses_.find(id);
    return it == masses_.end() ? 0.0 : it->second;
}

double Database::volume(size_t id) const {
    auto it = volumes_.find(id);
    if (it == volumes_.end() || it->second.v.empty()) throw Error("database: no size entry for element " + atom_label(id));
    const Fn &f = it->second;
    if (f.type == 0) return f.v[0];
    if (f.type == 1) return (4.0 / 3.0) * M_PI * powf((float)f.v[0], 3);
    if (f.type == 2) return std::sqrt(powf((float)M_PI, 3)) * powf((float)f.v[0], 3);
    throw Error("Size-type not implemented: type=" + std::to_string(f.type));
}

double Database::exclusionfactor(size_t id, double effvolume, double q) const {
    auto it = exclusion_.find(id);
    if (it == exclusion_.end() || it->second.v.empty())
        throw Error("database: no exclusionfactor entry for element " + atom_label(id));
    const Fn &f = it->second;
    if (f.type == 0) return f.v[0];
    if (f.type == 1) return effvolume * f.v[0];
    if (f.type == 2)
        return effvolume * std::exp(-1.0 * powf((float)effvolume, (float)(2.0 / 3.0)) * powf((float)q, 2) / (4 * M_PI)) * f.v[0];
    throw Error("ExclusionParameter-type not implemented: type=" + std::to_string(f.type));
}

double Database::sfactor(size_t id, double q) const {
    auto it = sfactors_.find(id);
    if (it == sfactors_.end() || it->second.v.empty())
        throw Error("database: no scatterfactor entry for element " + atom_label(id));
    const Fn &f = it->second;
    const std::vector<double> &v = f.v;
    if (f.type == 0) return v[0];
    if (f.type == 1) {  // Slater, 15 parameters (database.cpp:484-507)
        if (v.size() < 15) throw Error("database: Slater scatterfactor needs 15 parameters");
        double den = 0.0;
        for (size_t j = 0; j < 5; ++j) {
            const size_t j2 = 5 + j, j3 = 10 + j;
            if (v[j] != 0.0) {
                double c = 0.0;
                c += powf((float)(2.0 * v[j] / v[j3]), (float)(v[j3] + 0.5));
                double fac = 1.0;  // boost::math::factorial<double>(unsigned(2 v[j3])), exact in the table's range
                for (unsigned k = 2; k <= (unsigned)(2.0 * v[j3]); k++) fac *= (double)k;
                c /= std::sqrt(fac);
                den += v[j2] * powf((float)(c * powf((float)q, (float)(v[j3] - 1.0)) * std::exp(-1.0 * v[j] * q / v[j3] / 1.0)), 2);
            }
        }
        return den / (4.0 * M_PI);
    }
    if (f.type == 2) {  // four gaussians + constant (database.cpp:511-520)
        if (v.size() < 9) throw Error("database: gaussian scatterfactor needs 9 parameters");
        double den = 0.0;
        den += v[0] * std::exp(-v[1] * q * q);
        den += v[2] * std::exp(-v[3] * q * q);
        den += v[4] * std::exp(-v[5] * q * q);
        den += v[6] * std::exp(-v[7] * q * q);
        den += v[8];
        return den;
    }
    throw Error("ScatterFactor-type not implemented: type=" + std::to_string(f.type));
}

// ---------------------------------------------------------------------------------------------------------------
// Sample
// ---------------------------------------------------------------------------------------------------------------
std::vector<size_t> read_pdb_atoms(const std::string &filename, Database &db) {
    std::ifstream input(filename.c_str());
    if (input.fail()) throw Error("Couldn't open structure file: " + filename + " Typo in filename?");
    std::vector<size_t> ids;
    std::string line;
    while (getline(input, line)) {
        if (line.substr(0, 6) == "ATOM  ") {
            std::string pdbname = line.size() >= 16 ? line.substr(12, 4) : (line.size() > 12 ? line.substr(12) : "");
            ids.push_back(db.atom_id(db.pdb_name(trim(pdbname))));
        }
    }
    return ids;
}

static std::vector<size_t> read_pdb_selection(const std::string &filename, const std::string &selector,
                                              const std::string &expression) {
    if (selector != "beta" && selector != "segid") throw Error("PDB Atomselection only supports beta and segid at the moment");
    std::regex expr(expression);
    std::ifstream pdb(filename.c_str());
    if (pdb.fail()) throw Error("Couldn't open selection file: " + filename);
    std::vector<size_t> ids;
    size_t linecounter = 0;
    std::string line;
    while (getline(pdb, line)) {
        if (line.substr(0, 6) == "ATOM  ") {
            std::string value;
            if (selector == "beta") value = line.size() > 60 ? line.substr(60, 6) : "";
            else value = line.size() > 72 ? line.substr(72, 4) : "";
            if (std::regex_match(trim(value), expr)) ids.push_back(linecounter);
            linecounter++;
        }
    }
    return ids;
}

static std::map<std::string, std::vector<size_t>> read_ndx_selection(const std::string &filename, const std::string &selector,
                                                                     const std::string &expression) {
    if (selector != "name") throw Error("NDX Atomselection only supports selection by 'name' at the moment");
    std::regex expr(expression);
    std::ifstream ndx(filename.c_str());
    if (ndx.fail()) throw Error("Couldn't open selection file: " + filename);
    std::map<std::string, std::vector<size_t>> indexes;
    std::string line, name;
    while (getline(ndx, line)) {
        size_t pos = line.find("[");
        if (pos != std::string::npos) {
            size_t pos2 = line.find("]");
            if (pos2 == std::string::npos) throw Error("ndx file is missing closing bracket");
            // atomselection_reader.cpp:51-54 takes pos2 - pos characters after the '[', i.e. up to AND INCLUDING the ']', and keeps
            // the first whitespace-delimited token: "[ grpA ]" is named "grpA", but "[grpB]" is named "grpB]".  Kept as is: the
            // names are what a configuration refers to (pinned against the reference's reader in tests/test_control_plane.py).
            std::stringstream cs(line.substr(pos + 1, pos2 - pos));
            std::string cleanname;
            cs >> cleanname;
            name = trim(cleanname);
        } else if (!name.empty() && std::regex_match(name, expr)) {
            std::stringstream ls(line);
            size_t index = 0;
            while (ls >> index) indexes[name].push_back(index - 1);  // ndx files are 1-based
        }
    }
    return indexes;
}

void init_selections(const Config &cfg, Database &db, LoadedSample &s, const std::string &) {
    const size_t natoms = s.atom_ids.size();
    for (auto &sp : cfg.selections) {
        if (sp.type == "index") {
            s.selections[sp.name] = sp.ids;
        } else if (sp.type == "range") {
            std::vector<size_t> ids;
            for (size_t i = sp.from; i <= sp.to; i++) ids.push_back(i);  // inclusive (atomselection.cpp:71-73)
            s.selections[sp.name] = ids;
        } else if (sp.type == "lexical") {
            // Atoms::select (atoms.cpp:129-144) matches the expression against the database label of every atom.  The
            // reference pushes ids_[i] (the element ID) instead of the atom index i; the evident intent — the indices
            // of the matching atoms — is implemented here.
            std::regex expr(sp.expression);
            std::vector<size_t> ids;
            for (size_t i = 0; i < natoms; ++i)
                if (std::regex_match(db.atom_label(s.atom_ids[i]), expr)) ids.push_back(i);
            s.selections[sp.name] = ids;
        } else if (sp.type == "file") {
            if (sp.format == "ndx") {
                for (auto &kv : read_ndx_selection(sp.filepath, sp.selector, sp.expression)) s.selections[kv.first] = kv.second;
            } else if (sp.format == "pdb") {
                s.selections[sp.name] = read_pdb_selection(sp.filepath, sp.selector, sp.expression);
            } else {
                throw Error("Selection file format not understood");
            }
        }
    }
    for (auto &kv : s.selections)
        for (size_t i : kv.second)
            if (i >= natoms) throw Error("selection " + kv.first + " refers to atom index " + std::to_string(i) + " beyond the structure");
    if (s.selections.count("system")) {  // reserved word (sample.cpp:88-92)
        s.selections["system_RENAMED_BY_SASSENA"] = s.selections["system"];
        s.selections.erase("system");
    }
    std::vector<size_t> all(natoms);
    for (size_t i = 0; i < natoms; i++) all[i] = i;
    s.selections["system"] = all;
    auto t = s.selections.find(cfg.stager_target);
    if (t == s.selections.end()) throw Error("stager.target selection not found: " + cfg.stager_target);
    s.target = t->second;
    if (s.target.empty()) throw Error("No atoms available. Aborting");
}

void load_frames(const Config &cfg, const Database &db, LoadedSample &s) {
    const size_t natoms = s.atom_ids.size();
    const size_t NT = s.target.size();
    s.frames.clear();
    s.NF = 0;
    // Phase 1 — the frame index: global frame number -> (frameset, frame within it), after first/last/stride and clones,
    // in the order of the configuration (Frames::add_frameset, frames.cpp:75-160).
    struct Source {
        std::shared_ptr<void> keep;                       // the open frameset
        std::function<void(size_t, float *)> read;        // frame `local` of that frameset -> xyz[natoms*3]
    };
    struct Entry {
        size_t source, local;
    };
    std::vector<Source> sources;
    std::vector<Entry> index;
    auto add_block = [&](Source src, size_t nframes, size_t clones) {
        sources.push_back(std::move(src));
        for (size_t c = 0; c < clones; c++)  // CloneFrameset: the same frames again (frames.cpp:61-67,860-872)
            for (size_t i = 0; i < nframes; i++) index.push_back({sources.size() - 1, i});
    };
    auto add_dcd = [&](const std::string &path, const SampleFramesetParameters &f, size_t clones) {
        auto fs = std::make_shared<DCDFrameset>(path);
        if (fs->number_of_atoms != natoms)
            throw Error("Atom number mismatch (dcd) " + std::to_string(fs->number_of_atoms) + " vs. (pdb) " + std::to_string(natoms));
        fs->trim_index(f.first, f.last, f.last_set, f.stride);
        Source src;
        src.keep = fs;
        src.read = [fs](size_t i, float *xyz) { fs->read_frame(i, xyz); };
        add_block(std::move(src), fs->number_of_frames, clones);
    };
    // PDBFrameset (frames.cpp:442-577): frames end at lines starting with "END" (END / ENDMDL), a trailing unterminated
    // frame counts, coordinates are columns 31-38 / 39-46 / 47-54 of the ATOM records.  A frame without ATOM records (the
    // "ENDMDL" + "END" tail of multi-model files) is dropped here; the reference would index it as a 0-atom frame.
    auto add_pdb = [&](const std::string &path, const SampleFramesetParameters &f, size_t clones) {
        std::ifstream in(path.c_str());
        if (in.fail()) throw Error("Couldn't open frameset file: " + path);
        std::vector<std::vector<float>> all;
        std::vector<float> cur;
        std::string line;
        auto close_frame = [&]() {
            if (!cur.empty()) {
                if (cur.size() != natoms * 3)
                    throw Error("Atom number mismatch (pdb frameset) " + std::to_string(cur.size() / 3) + " vs. (structure) " +
                                std::to_string(natoms));
                all.push_back(cur);
            }
            cur.clear();
        };
        while (getline(in, line)) {
            if (line.compare(0, 6, "ATOM  ") == 0) {
                if (line.size() < 54) throw Error("short ATOM record in " + path);
                cur.push_back((float)atof(line.substr(30, 8).c_str()));
                cur.push_back((float)atof(line.substr(38, 8).c_str()));
                cur.push_back((float)atof(line.substr(46, 8).c_str()));
            } else if (line.compare(0, 3, "END") == 0) {
                close_frame();
            }
        }
        close_frame();
        auto kept = std::make_shared<std::vector<std::vector<float>>>();
        for (size_t i = 0; i < all.size(); i++) {  // FileFrameset::trim_index (frames.cpp:224-245)
            if (i < f.first || (f.last_set && i > f.last) || (f.stride > 1 && i % f.stride != 0)) continue;
            kept->push_back(std::move(all[i]));
        }
        Source src;
        src.keep = kept;
        src.read = [kept](size_t i, float *xyz) { std::copy((*kept)[i].begin(), (*kept)[i].end(), xyz); };
        add_block(std::move(src), kept->size(), clones);
    };
    // XTCFrameset / TRRFrameset (frames.cpp:592-858).  The reference insists on a pre-built .tnx frame index for these
    // formats (frames.cpp:133-139); the index is rebuilt in memory here, so none is needed.
    auto add_xdr = [&](std::shared_ptr<XdrFrameset> fs, const char *what, const SampleFramesetParameters &f, size_t clones) {
        if (fs->number_of_atoms != natoms)
            throw Error(std::string("Atom number mismatch (") + what + ") " + std::to_string(fs->number_of_atoms) + " vs. (pdb) " +
                        std::to_string(natoms));
        fs->trim_index(f.first, f.last, f.last_set, f.stride);
        Source src;
        src.keep = fs;
        src.read = [fs](size_t i, float *xyz) { fs->read_frame(i, xyz); };
        add_block(std::move(src), fs->number_of_frames, clones);
    };
    for (auto &f : cfg.framesets) {
        if (f.clones == 0) continue;
        if (f.format == "pdb") {
            add_pdb(f.filepath, f, f.clones);
        } else if (f.format == "pdblist") {  // frames.cpp:113-126
            std::ifstream list(f.filepath.c_str());
            if (list.fail()) throw Error("Couldn't open pdblist file: " + f.filepath);
            std::string line;
            while (list >> line) {
                if (!line.empty() && line[0] == '#') continue;
                for (size_t c = 0; c < f.clones; c++) add_pdb(cfg.get_filepath(line), f, 1);
            }
        } else if (f.format == "xtc") {
            add_xdr(std::make_shared<XTCFrameset>(f.filepath), "xtc", f, f.clones);
        } else if (f.format == "trr") {
            add_xdr(std::make_shared<TRRFrameset>(f.filepath), "trr", f, f.clones);
        } else if (f.format == "dcd") {
            add_dcd(f.filepath, f, f.clones);
        } else if (f.format == "dcdlist") {
            std::ifstream list(f.filepath.c_str());
            if (list.fail()) throw Error("Couldn't open dcdlist file: " + f.filepath);
            std::string line;
            while (list >> line) {
                if (!line.empty() && line[0] == '#') continue;
                for (size_t c = 0; c < f.clones; c++) add_dcd(cfg.get_filepath(line), f, 1);
            }
        } else {
            throw Error("frameset format '" + f.format + "' is not supported (dcd, dcdlist, pdb, pdblist, xtc, trr are)");
        }
    }
    const size_t NF = index.size();
    if (NF < 1) throw Error("No frames available. Aborting");
    auto load_raw = [&](size_t g, float *xyz) { sources[index[g].source].read(index[g].local, xyz); };
    // Phase 2 — alignments and motions (CoordinateSets::init, coordinate_sets.cpp:58-243): references are read now
    CoordinateSetsProcessor proc(cfg, db, s, NF, load_raw);
    // Phase 3 — CoordinateSets::load per frame (coordinate_sets.cpp:245-357) and the stager's narrowing to float
    std::vector<float> buf(natoms * 3);
    std::vector<double> work;
    s.frames.reserve(NF * NT * 3);
    for (size_t g = 0; g < NF; g++) {
        load_raw(g, buf.data());
        if (proc.active()) {
            work.assign(buf.begin(), buf.end());
            proc.apply(g, work);
            for (size_t a = 0; a < NT; a++)
                for (int c = 0; c < 3; c++) s.frames.push_back((float)work[3 * s.target[a] + c]);
        } else {
            for (size_t a = 0; a < NT; a++)
                for (int c = 0; c < 3; c++) s.frames.push_back(buf[3 * s.target[a] + c]);
        }
    }
    s.NF = NF;
}

ScatterFactors::ScatterFactors(const Config &cfg, const Database &db, const LoadedSample &s)
    : db_(db), sample_(s), kappas_(s.atom_ids.size(), 1.0), background_(cfg.background_factor) {
    for (auto &k : cfg.kappas) {  // scatter_factors.cpp:34-53
        auto it = s.selections.find(k.selection);
        if (it == s.selections.end()) throw Error("kappa selection not found: " + k.selection);
        for (size_t i : it->second) kappas_[i] = k.value;
    }
}

void ScatterFactors::update(double ql, double *factors) const {
    for (size_t i = 0; i < sample_.target.size(); ++i) {
        const size_t atom = sample_.target[i];
        const size_t id = sample_.atom_ids[atom];
        double sf = db_.sfactor(id, ql);
        if (background_ != 0.0) {
            const double k = kappas_[atom];
            const double v = db_.volume(id);
            const double efactor = db_.exclusionfactor(id, k * v, ql);
            sf = sf - background_ * efactor;
        }
        factors[i] = sf;
    }
}

double ScatterFactors::compute_background(double ql) const {
    double efactor_sum = 0, sf_sum = 0;
    for (size_t i = 0; i < sample_.target.size(); ++i) {
        const size_t atom = sample_.target[i];
        const size_t id = sample_.atom_ids[atom];
        sf_sum += db_.sfactor(id, ql);
        efactor_sum += db_.exclusionfactor(id, kappas_[atom] * db_.volume(id), ql);
    }
    return sf_sum / efactor_sum;
}

// ---------------------------------------------------------------------------------------------------------------
// output + driver
// ---------------------------------------------------------------------------------------------------------------
void write_npy(const std::string &path, const double *data, const std::vector<size_t> &shape) {
    std::string sh = "(";
    size_t n = 1;
    for (size_t i = 0; i < shape.size(); i++) {
        sh += std::to_string(shape[i]) + (shape.size() == 1 || i + 1 < shape.size() ? "," : "");
        if (i + 1 < shape.size()) sh += " ";
        n *= shape[i];
    }
    sh += ")";
    std::string hdr = "{'descr': '<f8', 'fortran_order': False, 'shape': " + sh + ", }";
    const size_t base = 10;  // magic(6) + version(2) + header length(2)
    size_t total = base + hdr.size() + 1;
    size_t pad = (64 - total % 64) % 64;
    hdr += std::string(pad, ' ');
    hdr += "\n";
    FILE *f = fopen(path.c_str(), "wb");
    if (!f) throw Error("cannot create " + path);
    const unsigned char magic[8] = {0x93, 'N', 'U', 'M', 'P', 'Y', 1, 0};
    fwrite(magic, 1, 8, f);
    uint16_t hl = (uint16_t)hdr.size();
    fwrite(&hl, 2, 1, f);
    fwrite(hdr.data(), 1, hdr.size(), f);
    if (n) fwrite(data, sizeof(double), n, f);
    fclose(f);
}

// reads back what write_npy wrote (little-endian float64, C order); false if the file does not exist
bool read_npy(const std::string &path, std::vector<double> &data, std::vector<size_t> &shape) {
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) return false;
    unsigned char pre[10];
    if (fread(pre, 1, 10, f) != 10 || memcmp(pre, "\x93NUMPY", 6) != 0) {
        fclose(f);
        throw Error("not an npy file: " + path);
    }
    const size_t hl = pre[8] | (pre[9] << 8);
    std::string hdr(hl, ' ');
    if (fread(&hdr[0], 1, hl, f) != hl || hdr.find("'<f8'") == std::string::npos || hdr.find("False") == std::string::npos) {
        fclose(f);
        throw Error("unsupported npy header in " + path);
    }
    shape.clear();
    size_t p = hdr.find("'shape'");
    p = hdr.find('(', p);
    const size_t e = hdr.find(')', p);
    size_t n = 1;
    for (size_t i = p + 1; i < e;) {
        while (i < e && !isdigit((unsigned char)hdr[i])) i++;
        if (i >= e) break;
        size_t v = 0;
        while (i < e && isdigit((unsigned char)hdr[i])) v = 10 * v + (hdr[i++] - '0');
        shape.push_back(v);
        n *= v;
    }
    data.resize(n);
    const bool ok = n == 0 || fread(data.data(), sizeof(double), n, f) == n;
    fclose(f);
    if (!ok) throw Error("short npy file: " + path);
    return true;
}

namespace {
// Row journal: every result row is appended to <dir>/rows.journal and flushed to the file system the moment the device hands
// it over, so a run that dies keeps every q-vector it has finished (the reference's HDF5WriterService appends rows as the
// partitions deliver them, file_writer_service.cpp:314-484, which is what makes its resume logic, sassena.cpp:270-305, useful).
// Record = q(3) | fq(2) | fq2(2) | fqt(2 NF) doubles, preceded once by a header {magic, NF}.
const double kJournalMagic = 20260117.5;

struct RowJournal {
    FILE *f = nullptr;
    std::string path;
    void open(const std::string &p, size_t NF) {
        path = p;
        f = fopen(p.c_str(), "wb");
        if (!f) throw Error("cannot create " + p);
        const double hdr[2] = {kJournalMagic, (double)NF};
        fwrite(hdr, sizeof(double), 2, f);
        fflush(f);
    }
    void append(const double q[3], const double *fqt, size_t NF, const double fq[2], const double fq2[2]) {
        if (!f) return;
        fwrite(q, sizeof(double), 3, f);
        fwrite(fq, sizeof(double), 2, f);
        fwrite(fq2, sizeof(double), 2, f);
        fwrite(fqt, sizeof(double), 2 * NF, f);
        fflush(f);
        fsync(fileno(f));
    }
    void close() {
        if (f) fclose(f);
        f = nullptr;
    }
    ~RowJournal() { close(); }
};

// complete records of a journal (a torn last record -- the process died inside append -- is dropped); false if absent
bool read_journal(const std::string &path, size_t NF, std::vector<double> &rows) {
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) return false;
    double hdr[2];
    rows.clear();
    if (fread(hdr, sizeof(double), 2, f) == 2 && hdr[0] == kJournalMagic && (size_t)hdr[1] == NF) {
        const size_t rec = 7 + 2 * NF;
        std::vector<double> r(rec);
        while (fread(r.data(), sizeof(double), rec, f) == rec) rows.insert(rows.end(), r.begin(), r.end());
    }
    fclose(f);
    return true;
}

// names in `dir` that start with `prefix`
std::vector<std::string> list_dir(const std::string &dir, const std::string &prefix) {
    std::vector<std::string> out;
    if (DIR *d = opendir(dir.c_str())) {
        while (dirent *e = readdir(d)) {
            const std::string n = e->d_name;
            if (n != "." && n != ".." && n.compare(0, prefix.size(), prefix) == 0) out.push_back(n);
        }
        closedir(d);
    }
    return out;
}

void remove_rank_dirs(const std::string &signal_dir) {
    for (const std::string &n : list_dir(signal_dir, "rank_")) {
        const std::string d = signal_dir + "/" + n;
        for (const std::string &fn : list_dir(d, "")) unlink((d + "/" + fn).c_str());
        rmdir(d.c_str());
    }
}

struct Collector : IResultSink {
    size_t NF = 0;
    std::vector<double> q, fqt, fq0, fq, fq2;
    RowJournal *journal = nullptr;
    SignalFileH5 *h5 = nullptr;  // single-rank HDF5 runs append to the file as they go and flush it periodically
    size_t flush_seconds = 600;
    std::chrono::steady_clock::time_point last_flush = std::chrono::steady_clock::now();
    void write(CartesianCoor3D qv, const double *t, size_t nf, std::complex<double> a, std::complex<double> a2) override {
        NF = nf;
        q.push_back(qv.x);
        q.push_back(qv.y);
        q.push_back(qv.z);
        fqt.insert(fqt.end(), t, t + 2 * nf);
        fq0.push_back(t[0]);
        fq0.push_back(t[1]);
        fq.push_back(a.real());
        fq.push_back(a.imag());
        fq2.push_back(a2.real());
        fq2.push_back(a2.imag());
        const double qq[3] = {qv.x, qv.y, qv.z}, f1[2] = {a.real(), a.imag()}, f2[2] = {a2.real(), a2.imag()};
        if (journal) journal->append(qq, t, nf, f1, f2);
        if (h5) {
            h5->write(qq, t, f1, f2);
            const auto now = std::chrono::steady_clock::now();
            if (std::chrono::duration<double>(now - last_flush).count() >= (double)flush_seconds) {
                h5->flush();  // file_writer_service.cpp:292-295: flush when serverflush seconds have passed
                last_flush = now;
            }
        }
    }
};
}  // namespace

// Params::overwrite_options (parameters.cpp:838-875).  The reference stores the three file names in `.file` only while every
// consumer reads `.filepath` (sample.cpp:36-37, data_stager.cpp:92-93, sassena.cpp:256), so there --sample.structure.file,
// --stager.file and --scattering.signal.file change nothing; here they take effect (the evident intent), resolved like the
// same element of the configuration file (get_filepath: relative to the configuration's directory, absolute paths kept).
static void overwrite_options(Config &cfg, const std::vector<std::pair<std::string, std::string>> &overwrites) {
    auto to_bool = [](const std::string &key, const std::string &v) {
        if (v == "1" || v == "true" || v == "on" || v == "yes") return true;
        if (v == "0" || v == "false" || v == "off" || v == "no") return false;
        throw Error("the argument ('" + v + "') for option '--" + key + "' is invalid");
    };
    for (const auto &kv : overwrites) {
        const std::string &key = kv.first, &val = kv.second;
        if (key == "sample.structure.file") {
            cfg.structure_file = val;
            cfg.structure_filepath = cfg.get_filepath(val);
        } else if (key == "sample.structure.format") {
            cfg.structure_format = val;
        } else if (key == "stager.target") {
            cfg.stager_target = val;
        } else if (key == "stager.dump") {
            cfg.stager.dump = to_bool(key, val);
        } else if (key == "stager.file") {
            cfg.stager.file = val;
            cfg.stager.filepath = cfg.get_filepath(val);
        } else if (key == "stager.format") {
            cfg.stager.format = val;
        } else if (key == "scattering.signal.file") {
            cfg.signal_file = val;
            cfg.signal_filepath = cfg.get_filepath(val);
        } else if (key == "limits.computation.threads") {
            size_t pos = 0;
            int n = 0;
            try {
                n = std::stoi(val, &pos);
            } catch (...) {
                pos = 0;
            }
            if (pos != val.size() || val.empty()) throw Error("the argument ('" + val + "') for option '--" + key + "' is invalid");
            (void)n;  // worker threads per process: no counterpart on the GPU path (accepted like the configuration's element)
        } else {
            throw Error("unrecognised option '--" + key + "'");
        }
    }
}

void Job::load(const std::string &config_file, const std::vector<std::pair<std::string, std::string>> &overwrites) {
    cfg.read_xml(config_file);
    overwrite_options(cfg, overwrites);
    db.read_xml(cfg.database_filepath);
    if (cfg.structure_format != "pdb") throw Error("structure format not supported: " + cfg.structure_format);
    sample.atom_ids = read_pdb_atoms(cfg.structure_filepath, db);
    init_selections(cfg, db, sample, cfg.structure_filepath);
    load_frames(cfg, db, sample);
    factors.reset(new ScatterFactors(cfg, db, sample));
}

size_t Job::run(const std::string &signal_dir_in, std::shared_ptr<ICommunicator> comm, const SgpuBackend &be, sgpu_ctx *ctx,
                std::string *report) {
    if (!factors) throw Error("Job::run before Job::load");
    LoadedSample &s = this->sample;
    const ScatterFactors &sf = *factors;
    Sample sample;
    sample.NA = s.target.size();
    sample.NF = s.NF;
    sample.frames = s.frames.data();
    sample.factors = [&](double ql, double *b) { sf.update(ql, b); };
    Collector sink;
    std::vector<CartesianCoor3D> qv = cfg.qvectors;
    // HDF5 output: <path>.h5 (+ the per-rank rows under <path>.h5.d/)
    const bool h5mode = signal_dir_in.size() > 3 && signal_dir_in.compare(signal_dir_in.size() - 3, 3, ".h5") == 0;
    const std::string signal_dir = h5mode ? signal_dir_in + ".d" : signal_dir_in;
    std::unique_ptr<SignalFileH5> h5file;
    bool h5_loaded = false;  // the writer holds the rows of an existing file
    if (h5mode) {
        // every rank reads the existing file (same file system) so that all agree on the q-vectors left to compute
        h5file.reset(new SignalFileH5(signal_dir_in, s.NF, cfg.signal_chunksize, cfg.signal_fqt, cfg.signal_fq0, cfg.signal_fq,
                                      cfg.signal_fq2));
        h5file->set_meta(cfg.rawconfig, cfg.rawconfig, db.rawconfig);
        // rows an interrupted run left in its journals go into the file first (rank 0), so that every rank then sees them as
        // done; stale per-rank directories of an earlier run are removed
        if (comm->rank() == 0) {
            std::vector<std::string> journals;
            for (const std::string &n : list_dir(signal_dir, "rank_")) journals.push_back(signal_dir + "/" + n + "/rows.journal");
            journals.push_back(signal_dir + "/rows.journal");
            bool opened = false;
            std::vector<double> have;
            for (const std::string &jp : journals) {
                std::vector<double> rows;
                if (!read_journal(jp, s.NF, rows) || rows.empty()) continue;
                if (!opened) {
                    have = h5file->init();
                    opened = true;
                }
                const size_t rec = 7 + 2 * s.NF;
                for (size_t i = 0; i + rec <= rows.size(); i += rec) {
                    const double *r = &rows[i];
                    bool done = false;
                    for (size_t k = 0; k + 2 < have.size() && !done; k += 3) done = have[k] == r[0] && have[k + 1] == r[1] && have[k + 2] == r[2];
                    if (done) continue;
                    h5file->write(r, r + 7, r + 3, r + 5);
                    have.insert(have.end(), r, r + 3);
                }
            }
            if (opened) h5file->flush();
            for (const std::string &jp : journals) unlink(jp.c_str());
            remove_rank_dirs(signal_dir);
        }
        comm->barrier();
        std::vector<double> old;  // (init() below re-reads the file, recovered rows included)
        {
            std::ifstream probe(signal_dir_in.c_str());
            if (probe.good()) {
                old = h5file->init();  // only rank 0 creates / rewrites the file (below)
                h5_loaded = true;
            }
        }
        if (!old.empty()) {  // only compute those q vectors which have not been written so far (sassena.cpp:277-291)
            std::vector<CartesianCoor3D> left;
            for (auto &q : qv) {
                bool done = false;
                for (size_t i = 0; i + 2 < old.size() && !done; i += 3) done = old[i] == q.x && old[i + 1] == q.y && old[i + 2] == q.z;
                if (!done) left.push_back(q);
            }
            qv.swap(left);
        }
    }
    // every writing rank (partition rank 0) journals its rows as they arrive; rows pair up through qvectors (arrival order,
    // as in the reference's HDF5 file, file_writer_service.cpp:314-484)
    mkdir(signal_dir.c_str(), 0777);
    if (!h5mode) {  // stale per-rank directories of an earlier run must not be mistaken for this run's rows
        if (comm->rank() == 0) remove_rank_dirs(signal_dir);
        comm->barrier();
    }
    std::string dir = signal_dir;
    if (comm->size() > 1) {
        dir = signal_dir + "/rank_" + std::to_string(comm->rank());
        mkdir(dir.c_str(), 0777);
    }
    RowJournal journal;
    journal.open(dir + "/rows.journal", s.NF);
    sink.journal = &journal;
    sink.flush_seconds = cfg.signal_flush_seconds;
    const bool h5direct = h5mode && comm->size() == 1;  // one rank: rows go straight into the HDF5 writer, flushed periodically
    if (h5direct) {
        if (!h5_loaded) h5file->init();
        sink.h5 = h5file.get();
    }
    std::unique_ptr<IScatterDevice> dev;
    if (!qv.empty()) dev.reset(ScatterDeviceFactory::create(comm, sample, &sink, qv, cfg, be, ctx));  // else: "No qvectors left to compute."
    if (dev) dev->run();
    journal.close();
    const size_t n = sink.q.size() / 3;
    if (n > 0 || comm->size() == 1 || h5mode) {  // (h5 mode always rewrites: stale rows of an earlier run must not be merged)
        write_npy(dir + "/qvectors.npy", sink.q.data(), {n, 3});
        if (cfg.signal_fqt) write_npy(dir + "/fqt.npy", sink.fqt.data(), {n, s.NF, 2});
        if (cfg.signal_fq0) write_npy(dir + "/fq0.npy", sink.fq0.data(), {n, 2});
        if (cfg.signal_fq) write_npy(dir + "/fq.npy", sink.fq.data(), {n, 2});
        if (cfg.signal_fq2) write_npy(dir + "/fq2.npy", sink.fq2.data(), {n, 2});
    }
    comm->barrier();
    if (h5direct) {
        h5file->flush();
        unlink(journal.path.c_str());
    } else if (h5mode && comm->rank() == 0) {
        // the reference's writer service runs on world rank 0 and receives the rows of every partition (file_writer_service
        // .cpp:252-310); here rank 0 collects them from the per-rank journals once all ranks are done
        if (!h5_loaded) h5file->init();
        const size_t rec = 7 + 2 * s.NF;
        for (size_t r = 0; r < comm->size(); r++) {
            const std::string jp = (comm->size() > 1 ? signal_dir + "/rank_" + std::to_string(r) : signal_dir) + "/rows.journal";
            std::vector<double> rows;
            if (!read_journal(jp, s.NF, rows)) continue;
            for (size_t i = 0; i + rec <= rows.size(); i += rec) h5file->write(&rows[i], &rows[i + 7], &rows[i + 3], &rows[i + 5]);
        }
        h5file->flush();
        for (size_t r = 0; r < comm->size(); r++)
            unlink(((comm->size() > 1 ? signal_dir + "/rank_" + std::to_string(r) : signal_dir) + "/rows.journal").c_str());
    } else if (!h5mode) {
        unlink(journal.path.c_str());  // the .npy datasets above hold the rows
    }
    comm->barrier();
    if (report) {
        std::ostringstream r;
        r << "atoms=" << s.atom_ids.size() << " target=" << s.target.size() << " frames=" << s.NF << " qvectors=" << cfg.qvectors.size()
          << " written=" << n << " background=" << (cfg.background_factor != 0.0 ? sf.compute_background(0.0) : 0.0);
        *report = r.str();
    }
    return n;
}

size_t Job::stage(std::shared_ptr<ICommunicator> comm, const SgpuBackend &be, sgpu_ctx *ctx, std::string *report) {
    if (!factors) throw Error("Job::stage before Job::load");
    LoadedSample &s = this->sample;
    Sample smp;
    smp.NA = s.target.size();
    smp.NF = s.NF;
    smp.frames = s.frames.data();
    Timer timer;
    timer.start("total");
    size_t bytes = 0;
    struct OwnedCtx {  // no context handed in: one on the device this process is bound to, for the duration of the call
        const SgpuBackend &be;
        sgpu_ctx *h = nullptr;
        ~OwnedCtx() {
            if (h) be.destroy(h);
        }
    } own{be};
    if (!ctx) {
        if (be.init(-1, &own.h)) throw Error(std::string("sgpu_init: ") + be.last_error(nullptr));
        ctx = own.h;
    }
    if (cfg.stager.mode == "frames") {
        DataStagerByFrame st(smp, *comm, *comm, timer, be, ctx, cfg);
        st.stage_block();
        bytes = DivAssignment(comm->size(), comm->rank(), smp.NF).size() * smp.NA * 3 * sizeof(float);
    } else if (cfg.stager.mode == "atoms") {
        DataStagerByAtom st(smp, *comm, *comm, timer, be, ctx, cfg);
        st.stage();
        bytes = ModAssignment(comm->size(), comm->rank(), smp.NA).size() * smp.NF * 3 * sizeof(float);
    } else {  // s_stage.cpp:228-231
        throw Error("Staging mode not understood stager.mode=" + cfg.stager.mode + "\nUse 'frames' or 'atoms'");
    }
    timer.stop("total");
    if (report) {
        std::ostringstream r;
        r << "stager.mode=" << cfg.stager.mode << " target=" << smp.NA << " frames=" << smp.NF << " staged_bytes=" << bytes;
        if (cfg.stager.dump) r << " dump=" << cfg.stager.filepath;
        for (const std::string &k : timer.keys()) r << " " << k << "=" << timer.sum(k);
        *report = r.str();
    }
    return bytes;
}

}  // namespace sassena
