// TEST INFRASTRUCTURE.  Include-path overlay used only to compile the reference's Thirdparty/DBoW2 sources where they lie
// (oracle/Makefile, target dbowref): the base cv:: stand-in plus the two things TemplatedVocabulary.h additionally expects
// from OpenCV's core header — the iostream/sstream includes it relies on transitively, and cv::FileStorage / cv::FileNode,
// which its virtual YAML save()/load() members name (TemplatedVocabulary.h:265-275, :1454-1622).  The YAML path is never
// executed by the oracle (vocabularies are loaded with loadFromTextFile, :1338-1423), so these are inert declarations.
#ifndef ORBGPU_DBOW_SHIM_CORE_HPP
#define ORBGPU_DBOW_SHIM_CORE_HPP
#include "../../../../shim/opencv2/core/core.hpp"
#include <fstream>
#include <iostream>
#include <sstream>
#include <stdexcept>
#include <string>

namespace cv {
class FileNode {
public:
    FileNode operator[](const std::string&) const { return FileNode(); }
    FileNode operator[](const char*) const { return FileNode(); }
    FileNode operator[](int) const { return FileNode(); }
    size_t size() const { return 0; }
    operator int() const { return 0; }
    operator float() const { return 0.f; }
    operator double() const { return 0.0; }
    operator std::string() const { return std::string(); }
};
class FileStorage {
public:
    enum { READ = 0, WRITE = 1 };
    FileStorage(const std::string&, int) {}
    bool isOpened() const { return false; }
    void release() {}
    FileNode operator[](const std::string&) const { return FileNode(); }
    FileNode operator[](const char*) const { return FileNode(); }
};
template <typename T> inline FileStorage& operator<<(FileStorage& fs, const T&) { return fs; }
}  // namespace cv
#endif
