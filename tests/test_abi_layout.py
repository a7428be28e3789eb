"""include/awfm_abi.h must have the reference's struct layouts (src/AwFmIndex.h): checked by compiling the same
probe against both headers where /root/reference exists, and by the header's own static assertions everywhere."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"

PROBE = r"""
#include HEADER
#include <stddef.h>
#include <stdio.h>
#define P(s) printf(#s " %zu\n", sizeof(struct s))
#define O(s,m) printf(#s "." #m " %zu\n", offsetof(struct s,m))
int main(void){
P(AwFmNucleotideBlock);P(AwFmAminoBlock);P(AwFmIndexConfiguration);P(AwFmCompressedSuffixArray);P(AwFmSearchRange);
P(AwFmIndex);P(AwFmKmerSearchData);P(AwFmKmerSearchList);
O(AwFmNucleotideBlock,baseOccurrences);O(AwFmAminoBlock,baseOccurrences);
O(AwFmIndexConfiguration,suffixArrayCompressionRatio);O(AwFmIndexConfiguration,kmerLengthInSeedTable);
O(AwFmIndexConfiguration,alphabetType);O(AwFmIndexConfiguration,keepSuffixArrayInMemory);O(AwFmIndexConfiguration,storeOriginalSequence);
O(AwFmCompressedSuffixArray,valueBitWidth);O(AwFmCompressedSuffixArray,values);O(AwFmCompressedSuffixArray,compressedByteLength);
O(AwFmIndex,versionNumber);O(AwFmIndex,featureFlags);O(AwFmIndex,bwtLength);O(AwFmIndex,bwtBlockList);O(AwFmIndex,prefixSums);
O(AwFmIndex,kmerSeedTable);O(AwFmIndex,fileHandle);O(AwFmIndex,config);O(AwFmIndex,fileDescriptor);O(AwFmIndex,suffixArrayFileOffset);
O(AwFmIndex,sequenceFileOffset);O(AwFmIndex,fastaVector);O(AwFmIndex,suffixArray);
O(AwFmKmerSearchData,kmerString);O(AwFmKmerSearchData,kmerLength);O(AwFmKmerSearchData,positionList);O(AwFmKmerSearchData,count);O(AwFmKmerSearchData,capacity);
O(AwFmKmerSearchList,capacity);O(AwFmKmerSearchList,count);O(AwFmKmerSearchList,kmerSearchData);
P(FastaVector);P(FastaVectorString);P(FastaVectorMetadata);P(FastaVectorMetadataVector);
O(FastaVector,sequence);O(FastaVector,header);O(FastaVector,metadata);O(FastaVectorMetadata,headerEndPosition);
O(FastaVectorMetadata,sequenceEndPosition);O(FastaVectorMetadataVector,data);O(FastaVectorMetadataVector,count);
O(FastaVectorString,charData);O(FastaVectorString,count);
printf("codes %d %d %d %d %d\n", AwFmSuccess, AwFmFileReadOkay, AwFmGeneralFailure, AwFmAllocationFailure, AwFmFileReadFail);
printf("alphabets %d %d %d\n", AwFmAlphabetAmino, AwFmAlphabetDna, AwFmAlphabetRna);
return 0;}
"""


def run_probe(tmp_path, header, includes):
    src = tmp_path / "probe.c"
    src.write_text(PROBE.replace("HEADER", f'"{header}"'))
    exe = tmp_path / ("probe_" + header.replace(".", "_"))
    subprocess.run(["/usr/bin/gcc", "-mavx2", *[f"-I{i}" for i in includes], str(src), "-o", str(exe)], check=True)
    return subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout


def test_mirror_header_compiles_with_its_static_asserts(tmp_path):
    out = run_probe(tmp_path, "awfm_abi.h", [os.path.join(ROOT, "include")])
    assert "AwFmIndex 112" in out and "AwFmKmerSearchData 32" in out


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "src")), reason="reference tree not present")
def test_layout_identical_to_reference_header(tmp_path):
    mine = run_probe(tmp_path, "awfm_abi.h", [os.path.join(ROOT, "include")])
    theirs = run_probe(tmp_path, "AwFmIndex.h", [os.path.join(REF, "src"), os.path.join(REF, "lib/FastaVector/src")])
    assert mine == theirs
